#include "oracle.h"
#include <stdio.h>
void orc_mhd3d_rotating_step(const orc_params *P, real_t *Uold, real_t *Unew, real_t dt, real_t t){(void)P;(void)Uold;(void)Unew;(void)dt;(void)t;}
void orc_make_all_boundaries_shear(const orc_params *P, real_t *U, real_t dt, real_t t){(void)P;(void)U;(void)dt;(void)t;}
