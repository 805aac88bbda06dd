// Host-side batched Update_RCONST_x (include/mistra_rconst.h).  The right-hand
// sides are generated from the mechanism tables (_gen/rconst_<x>.inc); the rate
// laws are hand-restated in rate_laws.h.
#include "../../include/mistra_rconst.h"
#include "rate_laws.h"

#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "rconst_common.h"

namespace {

using rconst_common::kDims;
using rconst_common::MechDims;
using rconst_common::spc_index;
using rconst_common::fill_indices;

void rc_g(const rate_ctx *cx, double *RC)
{
#include "_gen/rconst_g.inc"
}
void rc_a(const rate_ctx *cx, double *RC)
{
#include "_gen/rconst_a.inc"
}
void rc_t(const rate_ctx *cx, double *RC)
{
#include "_gen/rconst_t.inc"
}

}  // namespace

extern "C" int mistra_rconst_spc_index(int mech, const char *name)
{
  if (mech < 0 || mech > 2 || !name) return -1;
  return spc_index(mech, name);
}

extern "C" int mistra_rconst_update(int mech, const mistra_rate_inputs *in, double *rconst,
                                    int nthreads)
{
  if (mech < 0 || mech > 2 || !in || !rconst || in->ncell < 0) return -1;
  if (!in->cb1 || !in->scal || !in->ph_rat || !in->conc) return -1;
  const MechDims dm = kDims[mech];
  const int nspec = dm.nvar + dm.nfix;
  rate_ctx proto;
  memset(&proto, 0, sizeof proto);
  proto.nspec = nspec;
  proto.f32 = in->f32_literals ? 1 : 0;
  fill_indices(mech, &proto);
  const std::vector<double> zeros((size_t)nspec * 4, 0.0);
  if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for num_threads(nthreads) schedule(static)
#endif
  for (int64_t c = 0; c < in->ncell; ++c) {
    rate_ctx cx = proto;
    const double *cb = in->cb1 + 4 * c, *sc = in->scal + 13 * c;
    cx.aircc = cb[0]; cx.te = cb[1]; cx.h2oppm = cb[2]; cx.pk = cb[3];
    cx.conv1 = sc[0]; cx.xhal = sc[1]; cx.xiod = sc[2]; cx.xhet1 = sc[3]; cx.xhet2 = sc[4];
    for (int k = 0; k < 4; ++k) { cx.xliq[k] = sc[5 + k]; cx.cvv[k] = sc[9 + k]; }
    cx.ph_rat = in->ph_rat + (size_t)MISTRA_NPHRXN * c;
    cx.C = in->conc + (size_t)nspec * c;
    cx.FIX = cx.C + dm.nvar;
    cx.yhenry = in->yhenry ? in->yhenry + (size_t)nspec * c : zeros.data();
    cx.yxkmt = in->yxkmt ? in->yxkmt + (size_t)nspec * dm.nkc * c : zeros.data();
    cx.ykef = in->ykef ? in->ykef + (size_t)nspec * dm.nkc * c : zeros.data();
    cx.ykeb = in->ykeb ? in->ykeb + (size_t)nspec * dm.nkc * c : zeros.data();
    cx.yxkmtd = in->yxkmtd ? in->yxkmtd + (size_t)nspec * 2 * c : zeros.data();
    cx.yxeq = in->yxeq ? in->yxeq + (size_t)nspec * c : zeros.data();
    cx.ycw = in->ycw ? in->ycw + (size_t)dm.nkc * c : zeros.data();
    cx.ycwd = in->ycwd ? in->ycwd + (size_t)2 * c : zeros.data();
    double *RC = rconst + (size_t)dm.nreact * c;
    const rate_ctx *cxp = &cx;
    switch (mech) {
      case 0: rc_g(cxp, RC); break;
      case 1: rc_a(cxp, RC); break;
      default: rc_t(cxp, RC); break;
    }
  }
  return 0;
}


// ---- per-layer tables of the liq_parm chain on the host (include/mistra_liq.h)
#include "liq_host.inc"
