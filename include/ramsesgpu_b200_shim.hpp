// C++ classes with the reference's names (HydroRunBase / MHDRunBase / HydroRunGodunov /
// MHDRunGodunov, namespace hydroSimu) forwarding to the C ABI of ramsesgpu_b200.h, so that code
// written against the reference's operator surface (src/hydro/HydroRunBase.h:63-639,
// MHDRunGodunov.h:58-260) can switch by changing an include.  Header-only.
#pragma once
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "ramsesgpu_b200.h"

namespace hydroSimu {

typedef double real_t;  // scalars cross the ABI as double whatever the solver precision

// the parameter-file object handed to the run constructors (reference ConfigMap); lookups are
// answered by the library's own parser through a throw-away handle-free path: we keep the text.
class ConfigMap {
 public:
  explicit ConfigMap(const std::string& filename) {
    std::ifstream in(filename.c_str(), std::ios::in | std::ios::binary);
    if (!in.good()) throw std::runtime_error("cannot read parameter file " + filename);
    std::stringstream ss;
    ss << in.rdbuf();
    text_ = ss.str();
  }
  static ConfigMap fromText(const std::string& text) { ConfigMap c; c.text_ = text; return c; }
  const std::string& text() const { return text_; }
  // minimal getBool (same truth table as the reference, ConfigMap.cpp:65-87); section/name are
  // matched case-insensitively like INIReader::makeKey
  bool getBool(const std::string& section, const std::string& name, bool dflt) const {
    std::string v = lookup(section, name);
    if (v.empty()) return dflt;
    if (v == "1" || v == "yes" || v == "true" || v == "on") return true;
    if (v == "0" || v == "no" || v == "false" || v == "off") return false;
    return dflt;
  }

 private:
  ConfigMap() {}
  static std::string lower(std::string s) {
    for (size_t i = 0; i < s.size(); ++i) s[i] = (char)std::tolower((unsigned char)s[i]);
    return s;
  }
  static std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
  }
  std::string lookup(const std::string& section, const std::string& name) const {
    std::istringstream in(text_);
    std::string line, sec, found;
    while (std::getline(in, line)) {
      std::string t = trim(line);
      if (t.empty() || t[0] == '#' || t[0] == ';') continue;
      if (t[0] == '[') { size_t e = t.find(']'); if (e != std::string::npos) sec = lower(t.substr(1, e - 1)); continue; }
      size_t eq = t.find('=');
      if (eq == std::string::npos || sec != lower(section)) continue;
      if (lower(trim(t.substr(0, eq))) == lower(name)) {
        std::string v = t.substr(eq + 1);
        size_t c = v.find(" ;");
        if (c != std::string::npos) v = v.substr(0, c);
        found = trim(v);
      }
    }
    return found;
  }
  std::string text_;
};

class HydroRunBase {
 public:
  HydroRunBase(ConfigMap& cfg, bool fp32 = false) : h_(nullptr) {
    check(rg_create(cfg.text().c_str(), fp32 ? RG_FLAG_FP32 : 0, &h_));
    check(rg_get_layout(h_, &layout_));
  }
  virtual ~HydroRunBase() { rg_destroy(h_); }

  // ---- the reference's virtual surface -------------------------------------------------------
  virtual real_t compute_dt(int useU = 0) { double dt; check(rg_compute_dt(h_, useU, &dt)); return dt; }
  virtual void make_all_boundaries(int which = 0) { check(rg_make_all_boundaries(h_, which)); }
  virtual int init_simulation(const std::string problemName) {
    int n = 0; check(rg_init_simulation(h_, problemName.c_str(), &n)); return n;
  }
  virtual void godunov_unsplit(int nStep, real_t dt) { check(rg_godunov_unsplit(h_, nStep, dt)); }
  virtual void oneStepIntegration(int& nStep, real_t& t, real_t& dt) { check(rg_one_step(h_, &nStep, &t, &dt)); }
  virtual void start() { check(rg_run(h_)); }
  virtual void output(int nStep) { check(rg_output(h_, nStep)); }
  // getData(nStep): device pointer of the buffer holding step nStep (reference HydroRunBase.h:437)
  void* getData(int nStep = 0) { void* p = nullptr; check(rg_get_data_device(h_, nStep % 2, &p)); return p; }
  // copyGpuToCpu(nStep) + getDataHost(nStep): host copy, reference layout, ghosts included
  void copyGpuToCpu(int nStep, std::vector<char>& host) {
    host.resize(stateBytes());
    check(rg_copy_to_host(h_, nStep % 2, host.data(), host.size()));
  }
  size_t stateBytes() const {
    return (size_t)layout_.isize * layout_.jsize * layout_.ksize * layout_.nvar * layout_.real_bytes;
  }
  const rg_layout& layout() const { return layout_; }
  rg_handle handle() { return h_; }

 protected:
  static void check(int rc) { if (rc != RG_OK) throw std::runtime_error(rg_last_error()); }
  rg_handle h_;
  rg_layout layout_;
};

class MHDRunBase : public HydroRunBase {
 public:
  MHDRunBase(ConfigMap& cfg, bool fp32 = false) : HydroRunBase(cfg, fp32) {}
  virtual real_t compute_dt_mhd(int useU = 0) { return compute_dt(useU); }
};

class HydroRunGodunov : public HydroRunBase {
 public:
  HydroRunGodunov(ConfigMap& cfg, bool fp32 = false) : HydroRunBase(cfg, fp32) {}
};

class MHDRunGodunov : public MHDRunBase {
 public:
  MHDRunGodunov(ConfigMap& cfg, bool fp32 = false) : MHDRunBase(cfg, fp32) {}
};

}  // namespace hydroSimu
