// Shared by the host (rconst_host.cpp) and device (rconst_kernels.cu) versions of
// Update_RCONST_x: mechanism sizes and the species indices the heterogeneous rate laws
// look up by name (as mk_interface does, /root/reference/src/utils.f90:84-140).
#pragma once
#include "rate_laws.h"

#include <cstdio>
#include <cstring>

extern "C" const char *mistra_kpp_spc_name_impl(int mech, int i);

namespace rconst_common {

struct MechDims { int nvar, nfix, nreact, nkc; };
static const MechDims kDims[3] = {{102, 3, 331, 2}, {257, 5, 979, 2}, {417, 7, 1627, 4}};

inline int spc_index(int mech, const char *name)
{
  const int n = kDims[mech].nvar + kDims[mech].nfix;
  for (int i = 0; i < n; ++i)
    if (!strcmp(mistra_kpp_spc_name_impl(mech, i), name)) return i;
  return -1;
}

inline void fill_indices(int mech, rate_ctx *cx)
{
  const int nvar = kDims[mech].nvar;
  cx->i_HNO3 = spc_index(mech, "HNO3");
  cx->i_N2O5 = spc_index(mech, "N2O5");
  cx->i_NH3 = spc_index(mech, "NH3");
  cx->i_H2SO4 = spc_index(mech, "H2SO4");
  cx->i_ClNO3 = spc_index(mech, "ClNO3");
  cx->i_BrNO3 = spc_index(mech, "BrNO3");
  char nm[32];
  for (int k = 0; k < 4; ++k) {
    snprintf(nm, sizeof nm, "Clml%d", k + 1); cx->i_Clml[k] = spc_index(mech, nm);
    snprintf(nm, sizeof nm, "Brml%d", k + 1); cx->i_Brml[k] = spc_index(mech, nm);
    snprintf(nm, sizeof nm, "H2Ol%d", k + 1);
    int f = spc_index(mech, nm);
    cx->if_H2Ol[k] = f >= 0 ? f - nvar : -1;
    if (k < 2) {
      snprintf(nm, sizeof nm, "HNO3l%d", k + 1); cx->i_HNO3l[k] = spc_index(mech, nm);
      snprintf(nm, sizeof nm, "NO3ml%d", k + 1); cx->i_NO3ml[k] = spc_index(mech, nm);
    }
  }
}


}  // namespace rconst_common
