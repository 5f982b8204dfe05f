#include "output.h"

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

namespace {
std::string outputBase(const RunParams& rp, const Layout& L, std::string* rankTag) {
  std::string base = rp.outputDir;
  if (!base.empty() && base.back() != '/') base += "/";
  base += rp.outputPrefix;
  rankTag->clear();
  if (L.nranks > 1) {
    char buf[32];
    std::snprintf(buf, sizeof buf, "_rank%04d", L.rank);
    *rankTag = buf;
  }
  return base;
}
}  // namespace

std::string vtiPath(const RunParams& rp, const Layout& L, int nStep) {
  std::string rankTag;
  const std::string base = outputBase(rp, L, &rankTag);
  return base + rankTag + "_" + stepString(nStep) + ".vti";
}

void writeRestartMeta(const std::string& vti, const RestartMeta& m) {
  std::FILE* f = std::fopen((vti + ".meta").c_str(), "w");
  if (!f) return;
  // %a: exact hexadecimal floats, the decimal values are for the reader's eyes only
  std::fprintf(f, "nStep %d\ntotalTime %a\ndt %a\ndtNext %a\nrank %d\nnranks %d\nkOffset %d\n# totalTime = %.17g, dt = %.17g, dtNext = %.17g\n",
               m.nStep, m.totalTime, m.dt, m.dtNext, m.rank, m.nranks, m.kOffset, m.totalTime, m.dt, m.dtNext);
  std::fclose(f);
}

bool readRestartMeta(const std::string& vti, RestartMeta* m) {
  std::FILE* f = std::fopen((vti + ".meta").c_str(), "r");
  if (!f) return false;
  char key[64], val[128];
  int got = 0;
  while (std::fscanf(f, "%63s %127[^\n]", key, val) == 2) {
    if (!std::strcmp(key, "nStep")) { m->nStep = std::atoi(val); ++got; }
    else if (!std::strcmp(key, "totalTime")) { m->totalTime = std::strtod(val, nullptr); ++got; }
    else if (!std::strcmp(key, "dt")) { m->dt = std::strtod(val, nullptr); ++got; }
    else if (!std::strcmp(key, "dtNext")) { m->dtNext = std::strtod(val, nullptr); }
    else if (!std::strcmp(key, "rank")) { m->rank = std::atoi(val); }
    else if (!std::strcmp(key, "nranks")) { m->nranks = std::atoi(val); }
    else if (!std::strcmp(key, "kOffset")) { m->kOffset = std::atoi(val); }
  }
  std::fclose(f);
  return got == 3;
}

std::string restartSlabName(const std::string& name, int rank, int nranks) {
  if (nranks <= 1) return name;
  char tag[32];
  std::snprintf(tag, sizeof tag, "_rank%04d", rank);
  const size_t r = name.find("_rank");
  if (r != std::string::npos && r + 9 <= name.size()) {
    bool digits = true;
    for (size_t q = r + 5; q < r + 9; ++q) digits = digits && name[q] >= '0' && name[q] <= '9';
    if (digits) return name.substr(0, r) + tag + name.substr(r + 9);
  }
  // mono-domain name <prefix>_<step>.vti: the tag goes in front of the step suffix
  const size_t dot = name.rfind(".vti");
  size_t us = (dot == std::string::npos) ? std::string::npos : name.rfind('_', dot);
  if (us == std::string::npos) return name + tag;
  return name.substr(0, us) + tag + name.substr(us);
}

template <typename T>
bool readVti(const std::string& path, const Layout& L, T* U, bool* ghostIncluded, std::string* msg, bool* global) {
  std::ifstream in(path.c_str(), std::ios::in | std::ios::binary);
  if (!in) { if (msg) *msg = "cannot open restart file '" + path + "'"; return false; }
  std::string header, line;
  const std::string marker = "<AppendedData encoding=\"raw\">";
  bool found = false;
  while (std::getline(in, line)) {
    header += line + "\n";
    if (line.find(marker) != std::string::npos) { found = true; break; }
  }
  if (!found || in.get() != '_') { if (msg) *msg = "'" + path + "' is not a raw-appended .vti"; return false; }
  int e[6] = {0, 0, 0, 0, 0, 0};
  const size_t w = header.find("WholeExtent=\"");
  if (w == std::string::npos || std::sscanf(header.c_str() + w + 13, "%d %d %d %d %d %d", e, e + 1, e + 2, e + 3, e + 4, e + 5) != 6) {
    if (msg) *msg = "'" + path + "': no WholeExtent";
    return false;
  }
  const char* type = sizeof(T) == 8 ? "Float64" : "Float32";
  if (header.find(type) == std::string::npos) { if (msg) *msg = "'" + path + "': precision does not match the run"; return false; }
  const int nx = e[1] + 1, ny = e[3] + 1, nz = e[5] + 1;
  const int gw = L.ghostWidth, kd = (L.dim == 2) ? 1 : L.ksize;
  int g;
  bool wholeGrid = false;  // inner cells of the global grid: this rank takes its planes
  if (nx == L.isize && ny == L.jsize && nz == kd) g = 0;
  else if (nx == L.isize - 2 * gw && ny == L.jsize - 2 * gw && nz == ((L.dim == 2) ? 1 : L.ksize - 2 * gw)) g = gw;
  else if (L.nranks > 1 && nx == L.isize - 2 * gw && ny == L.jsize - 2 * gw && nz == L.nz) { g = gw; wholeGrid = true; }
  else { if (msg) *msg = "'" + path + "': grid size does not match the run"; return false; }
  if (ghostIncluded) *ghostIncluded = (g == 0);
  if (global) *global = wholeGrid;
  const int k0 = (L.dim == 2) ? 0 : g;
  const int kFirst = wholeGrid ? L.kOffset : 0, kCount = wholeGrid ? L.nzLocal : nz;
  const size_t plane = (size_t)L.isize * L.jsize, comp = plane * L.ksize, n = (size_t)nx * ny * nz;
  for (int v = 0; v < L.nvar; ++v) {
    uint32_t nbytes = 0;
    in.read(reinterpret_cast<char*>(&nbytes), sizeof nbytes);
    if (!in || nbytes != (uint32_t)(n * sizeof(T))) { if (msg) *msg = "'" + path + "': truncated or wrong variable count"; return false; }
    if (kFirst > 0) in.seekg((std::streamoff)kFirst * ny * nx * (std::streamoff)sizeof(T), std::ios::cur);
    for (int k = 0; k < kCount; ++k)
      for (int j = 0; j < ny; ++j) {
        T* dst = U + (size_t)v * comp + (size_t)(k + k0) * plane + (size_t)(j + g) * L.isize + g;
        in.read(reinterpret_cast<char*>(dst), (std::streamsize)nx * sizeof(T));
      }
    if (nz - kFirst - kCount > 0) in.seekg((std::streamoff)(nz - kFirst - kCount) * ny * nx * (std::streamoff)sizeof(T), std::ios::cur);
    if (!in) { if (msg) *msg = "'" + path + "': truncated"; return false; }
  }
  return true;
}

template <typename T>
void writeOutputs(const RunParams& rp, const Layout& L, const T* U, int nStep) {
  std::string rankTag;
  const std::string base = outputBase(rp, L, &rankTag);
  if (rp.outputVtk && !rp.outputVtkAscii)
    writeVti<T>(base + rankTag + "_" + stepString(nStep) + ".vti", L, U, rp.ghostIncluded);
  if (rp.outputXsm)
    writeXsm<T>(base + rankTag + "_" + std::string(kVarPrefix[ID]) + "_" + stepString(nStep) + ".xsm", L, U, ID);
}

template void writeVti<double>(const std::string&, const Layout&, const double*, bool);
template void writeVti<float>(const std::string&, const Layout&, const float*, bool);
template bool readVti<double>(const std::string&, const Layout&, double*, bool*, std::string*, bool*);
template bool readVti<float>(const std::string&, const Layout&, float*, bool*, std::string*, bool*);
template void writeXsm<double>(const std::string&, const Layout&, const double*, int);
template void writeXsm<float>(const std::string&, const Layout&, const float*, int);
template void writeOutputs<double>(const RunParams&, const Layout&, const double*, int);
template void writeOutputs<float>(const RunParams&, const Layout&, const float*, int);

}  // namespace rg
