#include "output.h"

#include <cstdint>
#include <cstdio>
#include <fstream>
#include <sstream>
#include <vector>

namespace rg {

namespace {
const char* kVarNames[8] = {"density", "energy", "mx", "my", "mz", "bx", "by", "bz"};
const char* kVarPrefix[8] = {"d", "p", "u", "v", "w", "a", "b", "c"};

std::string stepString(int nStep) {
  char buf[16];
  std::snprintf(buf, sizeof buf, "%07d", nStep);
  return buf;
}
}  // namespace

template <typename T>
void writeVti(const std::string& path, const Layout& L, const T* U, bool ghostIncluded) {
  const int g = ghostIncluded ? 0 : L.ghostWidth;
  const int nx = L.isize - 2 * g, ny = L.jsize - 2 * g;
  const int nz = (L.dim == 2) ? 1 : L.ksize - 2 * g;
  const int k0 = (L.dim == 2) ? 0 : g;
  const size_t plane = (size_t)L.isize * L.jsize, comp = plane * L.ksize;
  std::ofstream out(path.c_str(), std::ios::out | std::ios::binary);
  const char* type = sizeof(T) == 8 ? "Float64" : "Float32";
  out << "<VTKFile type=\"ImageData\" version=\"0.1\" byte_order=\"LittleEndian\">\n";
  std::ostringstream ext;
  ext << 0 << " " << nx - 1 << " " << 0 << " " << ny - 1 << " " << 0 << " " << nz - 1;
  out << "  <ImageData WholeExtent=\"" << ext.str() << "\" Origin=\"0 0 0\" Spacing=\"1 1 1\">\n";
  out << "  <Piece Extent=\"" << ext.str() << "\">\n";
  out << "    <PointData>\n";
  const size_t n = (size_t)nx * ny * nz;
  for (int v = 0; v < L.nvar; ++v) {
    const char* name = kVarNames[v];
    out << "     <DataArray type=\"" << type << "\" Name=\"" << name << "\" format=\"appended\" offset=\""
        << v * n * sizeof(T) + v * sizeof(uint32_t) << "\" />\n";
  }
  out << "    </PointData>\n    <CellData>\n    </CellData>\n  </Piece>\n  </ImageData>\n";
  out << "  <AppendedData encoding=\"raw\">\n_";
  std::vector<T> row(nx);
  const uint32_t nbytes = (uint32_t)(n * sizeof(T));
  for (int v = 0; v < L.nvar; ++v) {
    out.write(reinterpret_cast<const char*>(&nbytes), sizeof nbytes);
    for (int k = 0; k < nz; ++k)
      for (int j = 0; j < ny; ++j) {
        const T* src = U + (size_t)v * comp + (size_t)(k + k0) * plane + (size_t)(j + g) * L.isize + g;
        out.write(reinterpret_cast<const char*>(src), (std::streamsize)nx * sizeof(T));
      }
  }
  out << "  </AppendedData>\n</VTKFile>\n";
}

template <typename T>
void writeXsm(const std::string& path, const Layout& L, const T* U, int iVar) {
  if (iVar < 0 || iVar >= L.nvar) return;
  const int g = L.ghostWidth;
  const int nx = L.isize - 2 * g, ny = L.jsize - 2 * g, nz = (L.dim == 2) ? 1 : L.ksize - 2 * g;
  const size_t plane = (size_t)L.isize * L.jsize, comp = plane * L.ksize;
  std::ofstream out(path.c_str(), std::ios::out | std::ios::binary);
  if (L.dim == 2)
    out << "Binary 1 " << nx << "x" << ny << " " << nx * ny << "(" << sizeof(T) << " byte reals)\n";
  else
    out << "Binary 1 " << nx << "x" << ny << "x" << nz << " " << nx * ny * nz << "(" << sizeof(T) << " byte reals)\n";
  const int k0 = (L.dim == 2) ? 0 : g;
  for (int k = 0; k < nz; ++k)
    for (int j = 0; j < ny; ++j) {
      const T* src = U + (size_t)iVar * comp + (size_t)(k + k0) * plane + (size_t)(j + g) * L.isize + g;
      out.write(reinterpret_cast<const char*>(src), (std::streamsize)nx * sizeof(T));
    }
}

template <typename T>
void writeOutputs(const RunParams& rp, const Layout& L, const T* U, int nStep) {
  std::string base = rp.outputDir;
  if (!base.empty() && base.back() != '/') base += "/";
  base += rp.outputPrefix;
  std::string rankTag;
  if (L.nranks > 1) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "_rank%04d", L.rank);
    rankTag = buf;
  }
  if (rp.outputVtk && !rp.outputVtkAscii)
    writeVti<T>(base + rankTag + "_" + stepString(nStep) + ".vti", L, U, rp.ghostIncluded);
  if (rp.outputXsm)
    writeXsm<T>(base + rankTag + "_" + std::string(kVarPrefix[ID]) + "_" + stepString(nStep) + ".xsm", L, U, ID);
}

template void writeVti<double>(const std::string&, const Layout&, const double*, bool);
template void writeVti<float>(const std::string&, const Layout&, const float*, bool);
template void writeXsm<double>(const std::string&, const Layout&, const double*, int);
template void writeXsm<float>(const std::string&, const Layout&, const float*, int);
template void writeOutputs<double>(const RunParams&, const Layout&, const double*, int);
template void writeOutputs<float>(const RunParams&, const Layout&, const float*, int);

}  // namespace rg
