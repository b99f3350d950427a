// Minimal stand-in for the handful of Boost.StringAlgo calls the hla-mapper sources make
// (split / is_any_of / replace_all / erase_all / to_upper).  Test infrastructure only: it
// lets oracle/Makefile compile the UNMODIFIED reference sources where Boost is absent.
// Semantics kept: split() replaces the output container, does not compress adjacent
// delimiters, and yields N+1 tokens for N delimiters (so "a,b," -> {"a","b",""}).
#ifndef HLM_SHIM_BOOST_ALGORITHM_STRING_HPP
#define HLM_SHIM_BOOST_ALGORITHM_STRING_HPP
#include <algorithm>
#include <cctype>
#include <string>
#include <vector>
namespace boost {
namespace algorithm {
struct any_of_pred {
  std::string set;
  bool operator()(char c) const { return set.find(c) != std::string::npos; }
};
inline any_of_pred is_any_of(const std::string& s) { return any_of_pred{s}; }
inline any_of_pred is_any_of(const char* s) { return any_of_pred{std::string(s)}; }
template <class Seq, class Pred>
inline Seq& split(Seq& out, const std::string& in, Pred pred) {
  Seq tmp;
  std::string cur;
  for (char c : in) {
    if (pred(c)) {
      tmp.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  tmp.push_back(cur);
  out.swap(tmp);
  return out;
}
inline void replace_all(std::string& s, const std::string& from, const std::string& to) {
  if (from.empty()) return;
  std::string out;
  std::string::size_type pos = 0, hit;
  while ((hit = s.find(from, pos)) != std::string::npos) {
    out.append(s, pos, hit - pos);
    out.append(to);
    pos = hit + from.size();
  }
  out.append(s, pos, std::string::npos);
  s.swap(out);
}
inline void erase_all(std::string& s, const std::string& what) { replace_all(s, what, ""); }
inline void to_upper(std::string& s) {
  for (auto& c : s) c = (char)std::toupper((unsigned char)c);
}
}  // namespace algorithm
using algorithm::erase_all;
using algorithm::is_any_of;
using algorithm::replace_all;
using algorithm::split;
using algorithm::to_upper;
}  // namespace boost
#endif
