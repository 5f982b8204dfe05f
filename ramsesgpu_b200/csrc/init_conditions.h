// Host-side initial conditions for the problems of the benchmark configurations.  Arrays are the
// LOCAL z-slab [var][k][j][i] (ghosts included); every generator works from GLOBAL indices and a
// single global pseudo-random stream so that the result does not depend on the number of slabs
// (the reference's MPI build reseeds per rank and therefore does depend on it,
// HydroRunBaseMpi.cpp:10009-10015).
#pragma once
#include <string>
#include <vector>

#include "config_map.h"
#include "params.h"

namespace rg {

// returns false (and leaves U zeroed) when the problem name is unknown for this solver family,
// like the reference which prints a message and carries on (MHDRunBase.cpp:1338-1341)
template <typename T>
bool initProblem(const ConfigMap& cfg, const RunParams& rp, const KParams<T>& kp, const std::string& problem,
                 std::vector<T>& U, std::string* message);

}  // namespace rg
