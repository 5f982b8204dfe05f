// Output writers for the two formats used by the parity checks: raw-appended VTK ImageData (.vti,
// all variables) and Xsmurf (.xsm, density).  Byte layout follows the reference's hand-written
// writers (HydroRunBase.cpp:2520-2562 and :2877-3037) so that the same readers work on both.
#pragma once
#include <string>

#include "params.h"
#include "run.h"

namespace rg {

template <typename T>
void writeVti(const std::string& path, const Layout& L, const T* U, bool ghostIncluded);
template <typename T>
void writeXsm(const std::string& path, const Layout& L, const T* U, int iVar);
// writes what the [output] section asks for (outputVtk / outputXsm) for step nStep
template <typename T>
void writeOutputs(const RunParams& rp, const Layout& L, const T* U, int nStep);


// Checkpoint / resume.  The reference restarts from its HDF5 output ("total time" and
// "time step" attributes, HydroRunBase.cpp:5100-5110, MHDRunBase.cpp:1234-1282); HDF5 is not available
// here, so the restart input is the raw-appended .vti written by writeOutputs plus a small text sidecar
// `<same name>.meta` holding what the .vti cannot: time step, total time, dt and the next dt as exact hex floats (the next dt
// so that a resumed run repeats the uninterrupted one bit for bit whichever kernel reduced it).
struct RestartMeta {
  int nStep = 0;
  double totalTime = 0.0, dt = 0.0;
  double dtNext = 0.0;  // CFL step of the dumped state as the uninterrupted run will use it (0 = recompute)
  // decomposition of the run that wrote the dump (validated on read: a slab file only resumes the same slab)
  int rank = 0, nranks = 1, kOffset = 0;
};
void writeRestartMeta(const std::string& vtiPath, const RestartMeta& m);
bool readRestartMeta(const std::string& vtiPath, RestartMeta* m);
// path of the .vti of step nStep for this rank (same naming as writeOutputs)
std::string vtiPath(const RunParams& rp, const Layout& L, int nStep);
// name of the dump rank `rank` of `nranks` resumes from when the parameter file names `restartFilename`:
// the name itself for a single rank; otherwise the same name with this rank's `_rankNNNN` tag (replacing the
// tag of another rank, or inserted in front of the `_<step>.vti` suffix of a mono-domain name)
std::string restartSlabName(const std::string& restartFilename, int rank, int nranks);
// reads a .vti written by writeVti into the local array U ([var][k][j][i], ghosts included; cells outside
// the file's extent are left untouched).  The file holds either exactly this rank's slab (with or without
// ghosts) or the inner cells of the GLOBAL grid, of which planes [kOffset, kOffset + nzLocal) are taken
// (*global = true).  Returns false with a message when the file does not match.
template <typename T>
bool readVti(const std::string& path, const Layout& L, T* U, bool* ghostIncluded, std::string* msg, bool* global = nullptr);

}  // namespace rg
