#ifndef HLM_SHIM_BOOST_REPLACE_HPP
#define HLM_SHIM_BOOST_REPLACE_HPP
#include <boost/algorithm/string.hpp>
#endif
