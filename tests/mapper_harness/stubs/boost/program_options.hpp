// stand-in for <boost/program_options.hpp> (tests/mapper_harness): the option-description surface src/mapper.cc uses
// (typed values bound to variables, required / default_value, parse_command_line of "--name value" / "--name=value",
// variables_map lookup with as<T>()).
#ifndef MAPPER_HARNESS_BOOST_PROGRAM_OPTIONS_
#define MAPPER_HARNESS_BOOST_PROGRAM_OPTIONS_
#include <climits>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>
namespace boost { namespace program_options {
struct error : std::runtime_error { error(const std::string& m) : std::runtime_error(m) {} };
struct value_semantic {
  bool is_required = false, has_default = false, has_value = false;
  virtual ~value_semantic() {}
  virtual void parse(const std::string& s) = 0;
  virtual void apply_default() = 0;
  virtual const void* get(const std::type_info& ti) const = 0;
};
template <typename T> struct typed_value : value_semantic {
  T* target; T def; T cur;
  explicit typed_value(T* t) : target(t), def(), cur() {}
  typed_value* required() { is_required = true; return this; }
  typed_value* default_value(const T& v) { def = v; has_default = true; return this; }
  typed_value* default_value(const T& v, const std::string& /*textual*/) { return default_value(v); }
  void parse(const std::string& s) override { std::istringstream ss(s); ss >> std::boolalpha >> cur; if (ss.fail()) { std::istringstream s2(s); s2 >> cur; if (s2.fail()) throw error("invalid option value '" + s + "'"); } has_value = true; if (target) *target = cur; }
  void apply_default() override { if (!has_value && has_default) { cur = def; has_value = true; if (target) *target = cur; } }
  const void* get(const std::type_info& ti) const override { if (ti != typeid(T)) throw error("bad option type"); return &cur; }
};
template <> inline void typed_value<std::string>::parse(const std::string& s) { cur = s; has_value = true; if (target) *target = cur; }
template <typename T> typed_value<T>* value(T* t = nullptr) { return new typed_value<T>(t); }
struct option_description { std::string long_name, short_name, help; std::shared_ptr<value_semantic> sem; };
class options_description;
struct options_description_easy_init {
  options_description* owner;
  options_description_easy_init& operator()(const char* name, const char* help);
  options_description_easy_init& operator()(const char* name, value_semantic* sem, const char* help = "");
};
class options_description {
 public:
  std::string caption; std::vector<option_description> opts;
  explicit options_description(const std::string& c = "") : caption(c) {}
  options_description_easy_init add_options() { options_description_easy_init e; e.owner = this; return e; }
  void add(const char* name, value_semantic* sem, const char* help) {
    option_description d; std::string n(name); const size_t k = n.find(','); d.long_name = n.substr(0, k); if (k != std::string::npos) d.short_name = n.substr(k + 1);
    d.help = help ? help : ""; d.sem.reset(sem); opts.push_back(d);
  }
};
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, const char* help) { owner->add(name, nullptr, help); return *this; }
inline options_description_easy_init& options_description_easy_init::operator()(const char* name, value_semantic* sem, const char* help) { owner->add(name, sem, help); return *this; }
inline std::ostream& operator<<(std::ostream& os, const options_description& d) { os << d.caption << ":\n"; for (const auto& o : d.opts) os << "  --" << o.long_name << "  " << o.help << "\n"; return os; }
struct variable_value {
  std::shared_ptr<value_semantic> sem; bool flag = false;
  template <typename T> const T& as() const { if (!sem) throw error("option has no value"); return *static_cast<const T*>(sem->get(typeid(T))); }
  bool empty() const { return !flag && (!sem || !sem->has_value); }
};
struct parsed_options { const options_description* desc; std::vector<std::pair<std::string, std::string>> kv; };
inline parsed_options parse_command_line(int argc, const char* const* argv, const options_description& d) {
  parsed_options p; p.desc = &d;
  for (int i = 1; i < argc; ++i) {
    std::string a(argv[i]);
    if (a.compare(0, 2, "--") == 0) a = a.substr(2);
    else if (a.compare(0, 1, "-") == 0) { const std::string sh = a.substr(1); a.clear(); for (const auto& o : d.opts) if (o.short_name == sh) a = o.long_name; if (a.empty()) throw error("unknown option -" + sh); }
    else throw error("unexpected argument '" + a + "'");
    std::string val; bool has_val = false; const size_t eq = a.find('=');
    if (eq != std::string::npos) { val = a.substr(eq + 1); a = a.substr(0, eq); has_val = true; }
    const option_description* od = nullptr; for (const auto& o : d.opts) if (o.long_name == a) od = &o;
    if (!od) throw error("unknown option --" + a);
    if (od->sem && !has_val) { if (i + 1 >= argc) throw error("missing value for --" + a); val = argv[++i]; }
    p.kv.push_back(std::make_pair(a, val));
  }
  return p;
}
class variables_map {
 public:
  std::map<std::string, variable_value> m; const options_description* desc = nullptr;
  std::size_t count(const std::string& k) const { auto it = m.find(k); return it != m.end() && !it->second.empty() ? 1 : 0; }
  const variable_value& operator[](const std::string& k) const { static variable_value none; auto it = m.find(k); return it == m.end() ? none : it->second; }
  void notify() { if (!desc) return; for (const auto& o : desc->opts) if (o.sem && o.sem->is_required && !o.sem->has_value) throw error("the option '--" + o.long_name + "' is required but missing"); }
};
inline void store(const parsed_options& p, variables_map& vm) {
  vm.desc = p.desc;
  for (const auto& o : p.desc->opts) { variable_value v; v.sem = o.sem; vm.m[o.long_name] = v; }
  for (const auto& kv : p.kv) { variable_value& v = vm.m[kv.first]; if (v.sem) v.sem->parse(kv.second); else v.flag = true; }
  for (const auto& o : p.desc->opts) if (o.sem) o.sem->apply_default();
}
inline void notify(variables_map& vm) { vm.notify(); }
}}
#endif
