// stand-in for <boost/algorithm/string.hpp> (tests/mapper_harness)
#ifndef MAPPER_HARNESS_BOOST_ALGO_STRING_
#define MAPPER_HARNESS_BOOST_ALGO_STRING_
#include <algorithm>
#include <string>
#include <vector>
namespace boost {
struct is_any_of { std::string set; is_any_of(const std::string& s) : set(s) {} bool operator()(char c) const { return set.find(c) != std::string::npos; } };
enum token_compress_mode_type { token_compress_on, token_compress_off };
template <typename Pred> void split(std::vector<std::string>& out, const std::string& in, Pred pred, token_compress_mode_type m = token_compress_off) {
  out.clear(); std::string cur; bool last_sep = false;
  for (char c : in) { if (pred(c)) { if (!(m == token_compress_on && last_sep)) { out.push_back(cur); cur.clear(); } last_sep = true; } else { cur += c; last_sep = false; } }
  out.push_back(cur);
}
inline void trim(std::string& s) { const char* ws = " \t\r\n"; s.erase(0, s.find_first_not_of(ws)); const size_t e = s.find_last_not_of(ws); if (e != std::string::npos) s.erase(e + 1); else s.clear(); }
inline void to_lower(std::string& s) { std::transform(s.begin(), s.end(), s.begin(), ::tolower); }
inline void to_upper(std::string& s) { std::transform(s.begin(), s.end(), s.begin(), ::toupper); }
inline std::string to_upper_copy(std::string s) { to_upper(s); return s; }
inline std::string to_lower_copy(std::string s) { to_lower(s); return s; }
inline bool starts_with(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }
}
#endif
