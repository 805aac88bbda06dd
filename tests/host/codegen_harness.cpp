// Host compilation of the generated device code (strict arithmetic) for codegen checks.
#include <cmath>
#include <cstring>
#define __device__
#define __noinline__
#define __forceinline__ inline
#define __constant__
#define __restrict__
#define KPP_STRICT 1
#include MECH_HEADER
using namespace MECH_NS;
extern "C" {
int h_nslot() { return NSLOT; }
int h_sg() { return S_G; }
int h_sy() { return S_Y; }
int h_srct() { return S_RCT; }
int h_sfix() { return S_FIX; }
int h_sk1() { return S_K1; }
void h_set_coef(const double *c) { for (int i = 0; i < NCOEF; ++i) c_coef[i] = c[i]; }
int h_ncoef() { return NCOEF; }
const char *h_coef_lit(int i) { return coef_literals[i]; }
void h_decomp(double *w) { decomp(w); }
int h_jacprep(double *w, double ghinv) { return jacprep(w, ghinv); }
void h_fun0(double *w) { fun<0>(w, 0.0, 0.0); }
void h_solve1(double *w) { solve<1>(w, 0.0, 0.0, 0.0); }
}
