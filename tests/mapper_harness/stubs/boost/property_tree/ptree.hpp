// stand-in for <boost/property_tree/ptree.hpp> + ini_parser (tests/mapper_harness): flat "section.key" string map
#ifndef MAPPER_HARNESS_BOOST_PTREE_
#define MAPPER_HARNESS_BOOST_PTREE_
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost { namespace property_tree {
class ptree {
 public:
  std::map<std::string, std::string> kv;
  template <typename T> T get(const std::string& k) const { auto it = kv.find(k); if (it == kv.end()) throw std::runtime_error("no such node: " + k); std::istringstream ss(it->second); T t; ss >> t; return t; }
  template <typename T> T get(const std::string& k, const T& def) const { auto it = kv.find(k); if (it == kv.end()) return def; std::istringstream ss(it->second); T t; ss >> t; return t; }
  template <typename T> void put(const std::string& k, const T& v) { std::ostringstream ss; ss.precision(17); ss << v; kv[k] = ss.str(); }
};
template <> inline std::string ptree::get<std::string>(const std::string& k) const { auto it = kv.find(k); if (it == kv.end()) throw std::runtime_error("no such node: " + k); return it->second; }
namespace ini_parser {
inline void read_ini(const std::string& path, ptree& pt) {
  std::ifstream f(path.c_str()); if (!f) throw std::runtime_error("cannot open " + path);
  std::string line, sec;
  while (std::getline(f, line)) {
    if (line.empty() || line[0] == ';' || line[0] == '#') continue;
    if (line[0] == '[') { sec = line.substr(1, line.find(']') - 1); continue; }
    const size_t eq = line.find('='); if (eq == std::string::npos) continue;
    pt.kv[(sec.empty() ? "" : sec + ".") + line.substr(0, eq)] = line.substr(eq + 1);
  }
}
inline void write_ini(const std::string& path, const ptree& pt) {
  std::ofstream f(path.c_str()); std::string sec;
  for (const auto& e : pt.kv) { const size_t dot = e.first.find('.'); const std::string s = dot == std::string::npos ? "" : e.first.substr(0, dot); if (s != sec) { f << "[" << s << "]\n"; sec = s; } f << (dot == std::string::npos ? e.first : e.first.substr(dot + 1)) << "=" << e.second << "\n"; }
}
}
using ini_parser::read_ini; using ini_parser::write_ini;
}}
#endif
