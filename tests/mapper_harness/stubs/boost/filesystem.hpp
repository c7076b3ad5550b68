// stand-in for <boost/filesystem.hpp> (tests/mapper_harness)
#ifndef MAPPER_HARNESS_BOOST_FILESYSTEM_
#define MAPPER_HARNESS_BOOST_FILESYSTEM_
#include <string>
#include <sys/stat.h>
namespace boost { namespace filesystem {
struct path {
  std::string s; path() {} path(const std::string& p) : s(p) {} path(const char* p) : s(p) {}
  std::string string() const { return s; }
  const std::string& native() const { return s; }
  path filename() const { const size_t k = s.find_last_of('/'); return path(k == std::string::npos ? s : s.substr(k + 1)); }
  path stem() const { const std::string f = filename().s; const size_t k = f.find_last_of('.'); return path(k == std::string::npos ? f : f.substr(0, k)); }
  path extension() const { const std::string f = filename().s; const size_t k = f.find_last_of('.'); return path(k == std::string::npos ? "" : f.substr(k)); }
  path parent_path() const { const size_t k = s.find_last_of('/'); return path(k == std::string::npos ? "" : s.substr(0, k)); }
  path operator/(const path& o) const { return path(s + "/" + o.s); }
};
inline bool exists(const path& p) { struct stat st; return stat(p.s.c_str(), &st) == 0; }
inline bool is_directory(const path& p) { struct stat st; return stat(p.s.c_str(), &st) == 0 && S_ISDIR(st.st_mode); }
inline bool create_directory(const path& p) { return mkdir(p.s.c_str(), 0777) == 0; }
inline bool create_directories(const path& p) { return mkdir(p.s.c_str(), 0777) == 0; }
inline path unique_path(const path& model = path("%%%%-%%%%-%%%%-%%%%")) { static unsigned long ctr = 0; std::string s = model.s; for (char& c : s) if (c == '%') c = "0123456789abcdef"[(ctr = ctr * 6364136223846793005ul + 1442695040888963407ul) >> 60]; return path(s); }
inline path temp_directory_path() { return path("/tmp"); }
inline bool remove(const path& p) { return ::remove(p.s.c_str()) == 0; }
}}
#endif
