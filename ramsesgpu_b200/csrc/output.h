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

}  // namespace rg
