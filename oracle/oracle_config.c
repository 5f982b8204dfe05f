/*
 * oracle_config.c -- TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * ini text -> orc_params with the reference's semantics:
 *   - inih line rules (src/utils/config/inih/ini.cpp:71-140): 200-char lines, '#'/';' full-line
 *     comments, " ;" inline comments, keys lower-cased "section.name"
 *     (inih/INIReader.cpp:94-101), later duplicates overwrite.
 *   - ConfigMap::getFloat parses with strtof -> FLOAT, even in the double build
 *     (src/utils/config/ConfigMap.cpp:41-49); getBool accepts 1/yes/true/on (…:65-87).
 *   - defaults live at the call sites of HydroParameters.h:196-330.
 */
#include "oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MAXKV 512
typedef struct { char key[128]; char val[200]; } kv_t;
typedef struct { kv_t kv[MAXKV]; int n; } cfg_t;

static void cfg_set(cfg_t *c, const char *sec, const char *name, const char *val) {
  char key[128];
  snprintf(key, sizeof key, "%s.%s", sec, name);
  for (char *s = key; *s; ++s) *s = (char)tolower((unsigned char)*s);
  for (int i = 0; i < c->n; ++i)
    if (!strcmp(c->kv[i].key, key)) { snprintf(c->kv[i].val, sizeof c->kv[i].val, "%s", val); return; }
  if (c->n < MAXKV) {
    snprintf(c->kv[c->n].key, sizeof c->kv[c->n].key, "%s", key);
    snprintf(c->kv[c->n].val, sizeof c->kv[c->n].val, "%s", val);
    c->n++;
  }
}
static const char *cfg_get(const cfg_t *c, const char *sec, const char *name) {
  char key[128];
  snprintf(key, sizeof key, "%s.%s", sec, name);
  for (char *s = key; *s; ++s) *s = (char)tolower((unsigned char)*s);
  for (int i = 0; i < c->n; ++i)
    if (!strcmp(c->kv[i].key, key)) return c->kv[i].val;
  return "";
}
static char *rstrip_(char *s) {
  char *p = s + strlen(s);
  while (p > s && isspace((unsigned char)*--p)) *p = '\0';
  return s;
}
static char *lskip_(char *s) { while (*s && isspace((unsigned char)*s)) s++; return s; }
/* ini.cpp:44-52: stop at c or at ';' preceded by whitespace */
static char *find_char_or_comment_(char *s, char c) {
  int was_ws = 0;
  while (*s && *s != c && !(was_ws && *s == ';')) { was_ws = isspace((unsigned char)*s); s++; }
  return s;
}
static void cfg_parse(cfg_t *c, const char *text) {
  char section[50] = "", prev[50] = "";
  const char *p = text;
  c->n = 0;
  while (*p) {
    char line[200];
    size_t n = 0;
    /* fgets(line, 200) semantics: at most 199 chars, rest spills into the next "line" */
    while (*p && *p != '\n' && n < sizeof line - 1) line[n++] = *p++;
    if (*p == '\n' && n < sizeof line - 1) p++;
    line[n] = '\0';
    char *start = lskip_(rstrip_(line));
    if (*prev && *start && start > line) {       /* multiline continuation, ini.cpp:96-102 */
      cfg_set(c, section, prev, start);
    } else if (*start == ';' || *start == '#') {
    } else if (*start == '[') {
      char *end = find_char_or_comment_(start + 1, ']');
      if (*end == ']') { *end = '\0'; snprintf(section, sizeof section, "%s", start + 1); *prev = '\0'; }
    } else if (*start) {
      char *end = find_char_or_comment_(start, '=');
      if (*end == '=') {
        *end = '\0';
        char *name = rstrip_(start);
        char *value = lskip_(end + 1);
        end = find_char_or_comment_(value, '\0');
        if (*end == ';') *end = '\0';
        rstrip_(value);
        snprintf(prev, sizeof prev, "%s", name);
        cfg_set(c, section, name, value);
      }
    }
  }
}
static long get_int(const cfg_t *c, const char *s, const char *n, long d) {
  const char *v = cfg_get(c, s, n); char *e; long r = strtol(v, &e, 0); return e > v ? r : d;
}
static float get_float(const cfg_t *c, const char *s, const char *n, float d) {
  const char *v = cfg_get(c, s, n); char *e; float r = strtof(v, &e); return e > v ? r : d;
}
static int get_bool(const cfg_t *c, const char *s, const char *n, int d) {
  const char *v = cfg_get(c, s, n); int r = d;
  if (!strcmp(v, "1") || !strcmp(v, "yes") || !strcmp(v, "true") || !strcmp(v, "on")) r = 1;
  if (!strcmp(v, "0") || !strcmp(v, "no") || !strcmp(v, "false") || !strcmp(v, "off")) r = 0;
  if (!*v) r = d;
  return r;
}
static void lower_copy(char *dst, size_t n, const char *src) {
  size_t i = 0; for (; src[i] && i + 1 < n; ++i) dst[i] = (char)tolower((unsigned char)src[i]); dst[i] = 0;
}

int orc_sizeof_real(void) { return (int)sizeof(real_t); }

long orc_array_len(const orc_params *p) {
  return (long)p->isize * p->jsize * p->ksize * p->nbVar;
}

int orc_params_from_ini(const char *text, orc_params *p) {
  static cfg_t c; /* big: keep off the stack */
  cfg_parse(&c, text);
  memset(p, 0, sizeof *p);

  /* HydroParameters.h:196-198 */
  p->nStepmax = (int)get_int(&c, "run", "nstepmax", 1000);
  p->tEnd = get_float(&c, "run", "tend", 0.0f);
  p->nOutput = (int)get_int(&c, "run", "noutput", 100);
  /* :201-210 */
  p->nx = (int)get_int(&c, "mesh", "nx", 2);
  p->ny = (int)get_int(&c, "mesh", "ny", 2);
  p->nz = (int)get_int(&c, "mesh", "nz", 1);
  p->dim = (p->nz == 1) ? 2 : 3;
  p->nbVar = (p->nz == 1) ? NVAR_2D : NVAR_3D;
  /* :230-235 */
  p->mhdEnabled = get_bool(&c, "MHD", "enable", 0);
  if (p->mhdEnabled) p->nbVar = NVAR_MHD;
  /* :238-247 */
  p->xMin = get_float(&c, "mesh", "xmin", 0.0f); p->xMax = get_float(&c, "mesh", "xmax", 1.0f);
  p->yMin = get_float(&c, "mesh", "ymin", 0.0f); p->yMax = get_float(&c, "mesh", "ymax", 1.0f);
  p->zMin = get_float(&c, "mesh", "zmin", 0.0f); p->zMax = get_float(&c, "mesh", "zmax", 1.0f);
  p->dx = (p->xMax - p->xMin) / p->nx;
  p->dy = (p->yMax - p->yMin) / p->ny;
  p->dz = (p->zMax - p->zMin) / p->nz;
  /* :253-258 */
  static const char *bcn[6] = {"boundary_xmin", "boundary_xmax", "boundary_ymin",
                               "boundary_ymax", "boundary_zmin", "boundary_zmax"};
  for (int f = 0; f < 6; ++f) p->bc[f] = (int)get_int(&c, "mesh", bcn[f], BC_DIRICHLET);
  /* :260-271 */
  p->ghostWidth = (int)get_int(&c, "mesh", "ghostWidth", 2);
  if (p->ghostWidth != 2 && p->ghostWidth != 3) p->ghostWidth = 2;
  if (p->mhdEnabled) p->ghostWidth = 3;
  /* :274-282 */
  p->cfl = get_float(&c, "hydro", "cfl", 0.5f);
  if (!p->cfl) p->cfl = 0.5;
  snprintf(p->problem, sizeof p->problem, "%s",
           *cfg_get(&c, "hydro", "problem") ? cfg_get(&c, "hydro", "problem") : "unknown");
  /* :292-330 */
  p->cIso = get_float(&c, "hydro", "cIso", 0.0f);
  p->gamma0 = get_float(&c, "hydro", "gamma0", 1.4f);
  p->smallr = get_float(&c, "hydro", "smallr", 1e-10f);
  p->smallc = get_float(&c, "hydro", "smallc", 1e-10f);
  p->niter_riemann = (int)get_int(&c, "hydro", "niter_riemann", 10);
  p->iorder = (int)get_int(&c, "hydro", "iorder", 2);
  p->smalle = (real_t)1e-7;
  p->smallp = p->smallc * p->smallc / p->gamma0;
  if (p->cIso > 0) p->smallp = p->smallr * p->cIso * p->cIso;
  p->smallpp = p->smallr * p->smallp;
  p->gamma6 = (p->gamma0 + 1.0f) / (2.0f * p->gamma0);
  p->Omega0 = get_float(&c, "MHD", "omega0", 0.0f);
  p->slope_type = get_float(&c, "hydro", "slope_type", 1.0f);
  if (get_int(&c, "hydro", "traceVersion", 1) == 0) p->slope_type = 0.0;
  p->nu = get_float(&c, "hydro", "nu", 0.0f);
  p->eta = get_float(&c, "MHD", "eta", 0.0f);
  /* :352-380 riemannSolver (hlld/llf only when MHD) */
  {
    char s[32]; lower_copy(s, sizeof s, *cfg_get(&c, "hydro", "riemannSolver") ? cfg_get(&c, "hydro", "riemannSolver") : "approx");
    p->riemannSolver = RS_APPROX;
    if (!strcmp(s, "hll")) p->riemannSolver = RS_HLL;
    else if (!strcmp(s, "hllc")) p->riemannSolver = RS_HLLC;
    else if (p->mhdEnabled && !strcmp(s, "hlld")) p->riemannSolver = RS_HLLD;
    else if (p->mhdEnabled && !strcmp(s, "llf")) p->riemannSolver = RS_LLF;
  }
  /* :388-417 */
  p->magRiemannSolver = MAG_HLLD;
  if (p->mhdEnabled) {
    char s[32]; lower_copy(s, sizeof s, *cfg_get(&c, "MHD", "magRiemannSolver") ? cfg_get(&c, "MHD", "magRiemannSolver") : "hlld");
    if (!strcmp(s, "hllf")) p->magRiemannSolver = MAG_HLLF;
    else if (!strcmp(s, "hlla")) p->magRiemannSolver = MAG_HLLA;
    else if (!strcmp(s, "roe")) p->magRiemannSolver = MAG_ROE;
    else if (!strcmp(s, "llf")) p->magRiemannSolver = MAG_LLF;
    else if (!strcmp(s, "upwind")) p->magRiemannSolver = MAG_UPWIND;
  }
  /* MHDRunGodunov.cpp:161-164 (2D -> 1, 3D -> 4); HydroRunGodunov.cpp:70-74 */
  p->implementationVersion = (int)get_int(&c, "MHD", "implementationVersion", p->dim == 2 ? 1 : 4);
  p->unsplitVersion = (int)get_int(&c, "hydro", "unsplitVersion", 1);
  /* HydroParameters.h:446-461 */
  p->isize = p->nx + 2 * p->ghostWidth;
  p->jsize = p->ny + 2 * p->ghostWidth;
  p->ksize = (p->nz == 1) ? 1 : p->nz + 2 * p->ghostWidth;

  /* problem blocks: MHDRunBase.cpp:1480-1487, :2708-2726; HydroRunBase.cpp:5456-5460, :5864-5892 */
  p->ot_direction = (int)get_int(&c, "OrszagTang", "direction", 0);
  if (p->ot_direction < 0 || p->ot_direction > 3) p->ot_direction = 0;
  p->ot_kt = get_float(&c, "OrszagTang", "kt", 0.0f);
  p->mri_density = get_float(&c, "MRI", "density", 1.0f);
  p->mri_beta = get_float(&c, "MRI", "beta", 400.0f);
  p->mri_amp = get_float(&c, "MRI", "amp", 0.01f);
  p->mri_densfluct = get_float(&c, "MRI", "density_fluctuations", 0.0f);
  p->mri_seed = (int)get_int(&c, "MRI", "seed", 0);
  snprintf(p->mri_type, sizeof p->mri_type, "%s",
           *cfg_get(&c, "MRI", "type") ? cfg_get(&c, "MRI", "type") : "noflux");
  p->implode_seed = (int)get_int(&c, "implode", "seed", 1);
  p->implode_amp = get_float(&c, "implode", "amplitude", 0.0f);
  p->kh_seed = (int)get_int(&c, "kelvin-helmholtz", "seed", 1);
  p->kh_amp = get_float(&c, "kelvin-helmholtz", "amplitude", 0.1f);
  p->kh_p_rand = get_bool(&c, "kelvin-helmholtz", "perturbation_rand", 1);
  p->kh_p_sine = get_bool(&c, "kelvin-helmholtz", "perturbation_sine", 0);
  p->kh_p_sine_robertson = get_bool(&c, "kelvin-helmholtz", "perturbation_sine_robertson", 0);
  p->kh_rho_in = get_float(&c, "kelvin-helmholtz", "rho_inner", 2.0f);
  p->kh_rho_out = get_float(&c, "kelvin-helmholtz", "rho_outer", 1.0f);
  p->kh_pressure = get_float(&c, "kelvin-helmholtz", "pressure", 2.5f);
  p->kh_inner = get_float(&c, "kelvin-helmholtz", "inner_size", 0.2f);
  p->kh_outer = get_float(&c, "kelvin-helmholtz", "outer_size", 0.2f);
  p->kh_vin = get_float(&c, "kelvin-helmholtz", "vflow_in", -0.5f);
  p->kh_vout = get_float(&c, "kelvin-helmholtz", "vflow_out", 0.5f);
  p->kh_mode = get_float(&c, "kelvin-helmholtz", "mode", 2.0f);
  p->kh_w0 = get_float(&c, "kelvin-helmholtz", "w0", 0.1f);
  p->kh_delta = get_float(&c, "kelvin-helmholtz", "delta", 0.03f);
  /* static gravity: HydroRunBase.cpp:253-260 (forced on for Rayleigh-Taylor), HydroParameters.h:322-324.
     Only init_hydro_Rayleigh_Taylor fills h_gravity with the static field (HydroRunBase.cpp:6400-6408);
     with any other problem the allocated array stays zero. */
  p->gravityEnabled = get_bool(&c, "gravity", "static", 0) || get_bool(&c, "gravity", "self", 0);
  if (!strcmp(p->problem, "Rayleigh-Taylor") || !strcmp(p->problem, "Keplerian-disk")) p->gravityEnabled = 1;
  p->gravity_x = p->gravity_y = p->gravity_z = 0;
  if (!strcmp(p->problem, "Rayleigh-Taylor") || !strcmp(p->problem, "falling-bubble")) {
    p->gravity_x = get_float(&c, "gravity", "static_field_x", 0.0f);
    p->gravity_y = get_float(&c, "gravity", "static_field_y", 0.0f);
    p->gravity_z = get_float(&c, "gravity", "static_field_z", 0.0f);
  }
  p->rt_amp = get_float(&c, "rayleigh-taylor", "amplitude", 0.01f);
  p->rt_d0 = get_float(&c, "rayleigh-taylor", "d0", 1.0f);
  p->rt_d1 = get_float(&c, "rayleigh-taylor", "d1", 2.0f);
  p->rt_random = get_bool(&c, "rayleigh-taylor", "randomEnabled", 0);
  p->rt_seed = (int)get_int(&c, "rayleigh-taylor", "random_seed", 33);
  p->rt_bx = get_float(&c, "rayleigh-taylor", "bx", 1e-8f);
  p->rt_by = get_float(&c, "rayleigh-taylor", "by", 1e-8f);
  p->rt_bz = get_float(&c, "rayleigh-taylor", "bz", 1e-8f);
  /* jet: HydroParameters.h:434-444 (enabled by the problem name only), MHDRunBase.cpp:1756-1758 */
  p->enableJet = !strcmp(p->problem, "jet");
  p->ijet = (int)get_int(&c, "jet", "ijet", 0);
  p->djet = get_float(&c, "jet", "djet", 1.0f);
  p->ujet = get_float(&c, "jet", "ujet", 0.0f);
  p->pjet = get_float(&c, "jet", "pjet", 0.0f);
#ifdef ORACLE_FLOAT
  p->cjet = sqrtf(p->gamma0 * p->pjet / p->djet);
#else
  p->cjet = sqrt(p->gamma0 * p->pjet / p->djet);
#endif
  p->offsetJet = (int)get_int(&c, "jet", "offsetJet", 0);
  p->jet_bx = get_float(&c, "jet", "BStatic_x", 0.0f);
  p->jet_by = get_float(&c, "jet", "BStatic_y", 0.0f);
  p->jet_bz = get_float(&c, "jet", "BStatic_z", 0.0f);
  /* gravity field: uniform for Rayleigh-Taylor, the vertical field of init_mhd_mri_grav_field for MRI with
     [gravity] static=yes (MHDRunBase.cpp:2763-2766), zero otherwise */
  p->mri_smoothGravity = get_bool(&c, "MRI", "smoothGravity", 0);
  p->mri_zFloor = get_float(&c, "MRI", "zFloor", 5.0f);
  p->mri_bcFloor = get_bool(&c, "MRI", "floor", 0);
  p->blast[0] = get_float(&c, "blast", "radius", (float)(0.25 * (p->xMax - p->xMin)));
  p->blast[1] = get_float(&c, "blast", "center_x", (float)((p->xMax + p->xMin) / 2));
  p->blast[2] = get_float(&c, "blast", "center_y", (float)((p->yMax + p->yMin) / 2));
  p->blast[3] = get_float(&c, "blast", "center_z", (float)((p->zMax + p->zMin) / 2));
  p->blast[4] = get_float(&c, "blast", "density_in", 1.0f);
  p->blast[5] = get_float(&c, "blast", "density_out", 1.0f);
  p->blast[6] = get_float(&c, "blast", "pressure_in", 10.0f);
  p->blast[7] = get_float(&c, "blast", "pressure_out", 0.1f);
  p->gresho[0] = get_float(&c, "Gresho_vortex", "center_x", (float)((p->xMax + p->xMin) / 2));
  p->gresho[1] = get_float(&c, "Gresho_vortex", "center_y", (float)((p->yMax + p->yMin) / 2));
  p->gresho[2] = get_float(&c, "Gresho_vortex", "v_bulk_x", 0.0f);
  p->gresho[3] = get_float(&c, "Gresho_vortex", "v_bulk_y", 0.0f);
  p->gresho[4] = get_float(&c, "Gresho_vortex", "v_bulk_z", 0.0f);
  p->riemann2d[0] = get_float(&c, "riemann2d", "x", 0.5f);
  p->riemann2d[1] = get_float(&c, "riemann2d", "y", 0.5f);
  p->riemannConfId = (int)get_int(&c, "hydro", "riemann_config_number", 0);
  p->kepler[0] = get_float(&c, "Keplerian-disk", "epsilon", 0.01f);
  p->kepler[1] = get_float(&c, "Keplerian-disk", "pressure", 1e-6f);
  p->kepler[2] = get_float(&c, "Keplerian-disk", "xCenter", (float)((p->xMax + p->xMin) / 2.0));
  p->kepler[3] = get_float(&c, "Keplerian-disk", "yCenter", (float)((p->yMax + p->yMin) / 2.0));
  p->kepler[4] = get_float(&c, "gravity", "g", 1.0f);
  p->bubble[0] = get_float(&c, "falling-bubble", "radius", 0.1f);
  p->bubble[1] = get_float(&c, "falling-bubble", "center_x", (float)((p->xMin + p->xMax) / 2));
  p->bubble[2] = get_float(&c, "falling-bubble", "center_y", (float)(p->yMin + 0.8 * (p->yMax - p->yMin)));
  p->bubble[3] = get_float(&c, "falling-bubble", "center_z", 0.0f);
  p->bubble[4] = get_float(&c, "falling-bubble", "v0", 0.0f);
  p->bubble[5] = get_float(&c, "falling-bubble", "d0", 2.0f);
  p->bubble[6] = get_float(&c, "falling-bubble", "d1", 1.0f);
  p->gravityMode = 0;
  if (p->gravityEnabled) {
    if (!strcmp(p->problem, "Rayleigh-Taylor") || !strcmp(p->problem, "falling-bubble")) p->gravityMode = 1;
    else if (!p->mhdEnabled && p->dim == 2 && !strcmp(p->problem, "Keplerian-disk")) p->gravityMode = 3;
    else if (p->mhdEnabled && (!strcmp(p->problem, "MRI") || !strcmp(p->problem, "Mri") || !strcmp(p->problem, "mri"))) p->gravityMode = 2;
  }
  return 0;
}
