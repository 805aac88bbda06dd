/* Boundary B1 (include/mistra_kpp_f77.h): per-cell replacements of
 * INTEGRATE_g/_a/_t (/root/reference/src/gas.f:710-773, aer.f:1408, tot.f:2812)
 * over the batched C ABI.
 *
 * The COMMON blocks /GDATA_g/, /GDATA_a/, /GDATA_t/ belong to the Fortran program.  The shim finds them at the
 * first call: an image registered with mistra_kpp_f77_bind() wins, otherwise the symbols gdata_g_ / gdata_a_ /
 * gdata_t_ are looked up in the process (dlsym over the global scope: a Fortran executable linked with
 * -rdynamic, or any shared object loaded before the call).  Nothing is assumed about the address of an undefined
 * weak symbol. */
#define _GNU_SOURCE
#include "../../include/mistra_kpp_f77.h"
#include "../../include/mistra_kpp.h"

#include <dlfcn.h>
#include <stdio.h>

static void *g_bound[3];

void mistra_kpp_f77_bind(int mech, void *gdata)
{
  if (mech >= 0 && mech < 3) g_bound[mech] = gdata;
}

static void *common_block(int mech, const char *sym)
{
  if (!g_bound[mech]) g_bound[mech] = dlsym(RTLD_DEFAULT, sym);
  if (!g_bound[mech])
    fprintf(stderr, " mistra_kpp: COMMON block %s not found (link the host with -rdynamic or call mistra_kpp_f77_bind)\n", sym);
  return g_bound[mech];
}

static void integrate_common(int mech, int nvar, double *C, double *RCONST, double *ATOL,
                             double *RTOL, double *STEPMIN, double *tin, double *tout)
{
  mistra_kpp_opts o;
  int32_t ierr = 0;
  double hexit = 0.0, texit = *tin;
  int i, rc;
  mistra_kpp_default_opts(&o);         /* IPAR(1)=0, IPAR(2)=1, IPAR(4)=2, RPAR(3)=1d-3 */
  for (i = 0; i < nvar; i++) {         /* gas.f:745-746 */
    RTOL[i] = 1.0e-3;
    ATOL[i] = 1.0e-25;
  }
  rc = mistra_kpp_integrate(mech, 1, RCONST, C + nvar, C, *tin, *tout, &o, &ierr, 0, &hexit,
                            &texit, 0);
  if (rc != 0) {
    fprintf(stderr, " mistra_kpp: %s\n", mistra_kpp_last_error());
    return;
  }
  if (ierr < 0)                         /* gas.f:764-767 */
    fprintf(stderr, " Rosenbrock: Unsucessful step at T=%g (IERR=%d)\n", *tin, (int)ierr);
  *tin = texit;                         /* gas.f:769 */
  *STEPMIN = hexit;                     /* gas.f:770 */
}

void integrate_g_(double *tin, double *tout)
{
  struct mistra_gdata_g *g = (struct mistra_gdata_g *)common_block(MISTRA_KPP_GAS, "gdata_g_");
  if (!g) return;
  integrate_common(MISTRA_KPP_GAS, 102, g->C, g->RCONST, g->ATOL, g->RTOL, &g->STEPMIN, tin, tout);
}
void integrate_a_(double *tin, double *tout)
{
  struct mistra_gdata_a *g = (struct mistra_gdata_a *)common_block(MISTRA_KPP_AER, "gdata_a_");
  if (!g) return;
  integrate_common(MISTRA_KPP_AER, 257, g->C, g->RCONST, g->ATOL, g->RTOL, &g->STEPMIN, tin, tout);
}
void integrate_t_(double *tin, double *tout)
{
  struct mistra_gdata_t *g = (struct mistra_gdata_t *)common_block(MISTRA_KPP_TOT, "gdata_t_");
  if (!g) return;
  integrate_common(MISTRA_KPP_TOT, 417, g->C, g->RCONST, g->ATOL, g->RTOL, &g->STEPMIN, tin, tout);
}
