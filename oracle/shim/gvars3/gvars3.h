// TEST INFRASTRUCTURE — stand-in for GVars3 (configuration values only): a typed name -> value
// registry.  Register() installs the default when the name is unknown, as GVars3 does with SILENT.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>
namespace GVars3 {
enum { SILENT = 1, HIDDEN = 2, FATAL_IF_NOT_DEFINED = 4 };
template <class T> inline std::map<std::string, std::shared_ptr<T>>& gv_store() { static std::map<std::string, std::shared_ptr<T>> s; return s; }
template <class T> inline T* gv_slot(const std::string& name, const T& def) {
  auto& s = gv_store<T>();
  auto it = s.find(name);
  if (it == s.end()) it = s.emplace(name, std::make_shared<T>(def)).first;
  return it->second.get();
}
template <class T> class gvar3 {
 public:
  gvar3() : p(nullptr) {}
  gvar3(const std::string& name, const T& def, int = 0) : p(gv_slot<T>(name, def)) {}
  T& operator*() { return *p; }
  const T& operator*() const { return *p; }
  T* operator->() { return p; }
  T* p;
};
template <> class gvar3<std::string> {
 public:
  gvar3() : p(nullptr) {}
  gvar3(const std::string& name, const std::string& def, int = 0) : p(gv_slot<std::string>(name, def)) {}
  std::string& operator*() { return *p; }
  std::string* operator->() { return p; }
  std::string* p;
};
struct GV3 {
  template <class T> static void Register(gvar3<T>& g, const std::string& name, const T& def, int = 0) { g.p = gv_slot<T>(name, def); }
  template <class T> static T& get(const std::string& name, const T& def = T(), int = 0) { return *gv_slot<T>(name, def); }
  template <class T> static void set(const std::string& name, const T& v) { *gv_slot<T>(name, v) = v; }
};
struct GV2T {
  template <class T> void Register(gvar3<T>& g, const std::string& name, const T& def, int = 0) { g.p = gv_slot<T>(name, def); }
  int GetInt(const std::string& name, int def = 0, int = 0) { return *gv_slot<int>(name, def); }
  double GetDouble(const std::string& name, double def = 0, int = 0) { return *gv_slot<double>(name, def); }
  std::string GetString(const std::string& name, const std::string& def = "", int = 0) { return *gv_slot<std::string>(name, def); }
};
static GV2T GV2;
typedef void (*GUICallbackProc)(void* ptr, std::string sCommand, std::string sParams);
struct GUIT {
  void RegisterCommand(const std::string&, GUICallbackProc, void*) {}
  void UnRegisterCommand(const std::string&) {}
  void UnRegisterAllCommands(void*) {}
  void ParseLine(const std::string&) {}
};
static GUIT GUI;
inline std::vector<std::string> ChopAndUnquoteString(const std::string& s) {
  std::vector<std::string> v; std::string cur;
  for (char c : s) { if (c == ' ') { if (!cur.empty()) v.push_back(cur); cur.clear(); } else if (c != '"') cur += c; }
  if (!cur.empty()) v.push_back(cur);
  return v;
}
}  // namespace GVars3
