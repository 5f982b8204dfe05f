// ramsesgpu_b200_main --param <file.ini>   : counterpart of the reference's src/euler_main.cpp
// (GetPot flags --param / --help only; the solver family is picked by [MHD] enable like
// euler_main.cpp:109), built on the C++ shim classes over the C ABI.
#include <cstdio>
#include <cstring>
#include <string>

#include "ramsesgpu_b200_shim.hpp"

int main(int argc, char** argv) {
  std::string param = "./jet.ini";  // same default as the reference
  bool fp32 = false;
  for (int a = 1; a < argc; ++a) {
    if (!std::strcmp(argv[a], "--param") && a + 1 < argc) param = argv[++a];
    else if (!std::strcmp(argv[a], "--fp32")) fp32 = true;
    else if (!std::strcmp(argv[a], "--help")) {
      std::printf("usage: %s --param <parameter file> [--fp32]\n", argv[0]);
      return 0;
    }
  }
  try {
    hydroSimu::ConfigMap configMap(param);
    const bool mhdEnabled = configMap.getBool("MHD", "enable", false);
    if (mhdEnabled) {
      hydroSimu::MHDRunGodunov run(configMap, fp32);
      run.start();
    } else {
      hydroSimu::HydroRunGodunov run(configMap, fp32);
      run.start();
    }
  } catch (const std::exception& e) {
    std::fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
  return 0;
}
