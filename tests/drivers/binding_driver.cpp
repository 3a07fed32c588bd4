// binding_driver.cpp — which library does a symbol resolve to in THIS process?  Reads mangled
// symbol names from stdin, looks each one up the way the dynamic linker does for a call from the
// main program (dlsym(RTLD_DEFAULT, ...)) and prints "<symbol> <basename of the defining object>".
// tests/test_abi.py runs it (a) linked with the drop-in in front of the host library and (b) linked
// against the host library alone with the drop-in LD_PRELOADed, and requires every hot-path symbol
// of SURVEY 8b to land in libmeep_b200*: a hot-path symbol that failed to interpose would silently
// run the reference's CPU loop.
#include <dlfcn.h>
#include <stdio.h>
#include <string.h>
#include <iostream>
#include <string>

#include "meep.hpp"

int main(int argc, char **argv) {
  meep::initialize mpi(argc, argv); // (forces the libraries to be loaded exactly as for a simulation)
  std::string name;
  while (std::getline(std::cin, name)) {
    if (name.empty()) continue;
    void *p = dlsym(RTLD_DEFAULT, name.c_str());
    Dl_info info;
    const char *lib = "UNRESOLVED";
    if (p && dladdr(p, &info) && info.dli_fname) {
      lib = strrchr(info.dli_fname, '/');
      lib = lib ? lib + 1 : info.dli_fname;
    }
    printf("%s %s\n", name.c_str(), lib);
  }
  return 0;
}
