#include "oracle.h"
#include <stdio.h>
real_t orc_compute_dt_hydro(const orc_params *P, const real_t *U){(void)P;(void)U;return 0;}
void orc_mhd2d_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt){(void)P;(void)Uold;(void)Unew;(void)dt;}
void orc_hydro_step_v1(const orc_params *P, const real_t *Uold, real_t *Unew, real_t dt){(void)P;(void)Uold;(void)Unew;(void)dt;}
void orc_mhd3d_rotating_step(const orc_params *P, real_t *Uold, real_t *Unew, real_t dt, real_t t){(void)P;(void)Uold;(void)Unew;(void)dt;(void)t;}
void orc_make_all_boundaries_shear(const orc_params *P, real_t *U, real_t dt, real_t t){(void)P;(void)U;(void)dt;(void)t;}
void orc_riemann_hydro(const orc_params *p, const real_t ql[5], const real_t qr[5], real_t flux[5]){(void)p;(void)ql;(void)qr;(void)flux;}
