#ifndef HLM_SHIM_BOOST_GZIP_HPP
#define HLM_SHIM_BOOST_GZIP_HPP
#include <boost/iostreams/filtering_stream.hpp>
#endif
