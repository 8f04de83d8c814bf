// Headless stand-in for libltdl, mapped onto dlopen with a ':'-separated search path.
// Test infrastructure only.
#pragma once
#include <dlfcn.h>
#include <string>
typedef void *lt_dlhandle;
inline std::string &stg_shim_ltpath() { static std::string p; return p; }
inline int lt_dlinit() { return 0; }
inline const char *lt_dlerror() { const char *e = dlerror(); return e ? e : ""; }
inline int lt_dlsetsearchpath(const char *p) { stg_shim_ltpath() = p ? p : ""; return 0; }
inline int lt_dladdsearchdir(const char *p) { stg_shim_ltpath() += std::string(":") + p; return 0; }
inline const char *lt_dlgetsearchpath() { return stg_shim_ltpath().c_str(); }
inline lt_dlhandle lt_dlopenext(const char *name) {
  std::string path = stg_shim_ltpath() + ":";
  size_t s = 0, e;
  while ((e = path.find(':', s)) != std::string::npos) {
    std::string dir = path.substr(s, e - s);
    s = e + 1;
    if (dir.empty()) continue;
    if (void *h = dlopen((dir + "/" + name + ".so").c_str(), RTLD_NOW | RTLD_GLOBAL)) return h;
  }
  return dlopen((std::string(name) + ".so").c_str(), RTLD_NOW | RTLD_GLOBAL);
}
inline void *lt_dlsym(lt_dlhandle h, const char *s) { return dlsym(h, s); }
