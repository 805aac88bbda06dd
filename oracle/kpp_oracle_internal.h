/* CPU ORACLE - TEST INFRASTRUCTURE ONLY (see kpp_oracle.h). */
#ifndef KPP_ORACLE_INTERNAL_H
#define KPP_ORACLE_INTERNAL_H

typedef struct kpp_mech_s {
  const char *name;
  int nvar, nfix, nreact, lu_nonzero;
  const int *lu_icol, *lu_crow, *lu_diag;   /* 0-based images of /SDATA_x/ */
  const char *const *spc_names;
  void (*fun)(const double *V, const double *F, const double *RCT, double *Vdot);
  void (*jac)(const double *V, const double *F, const double *RCT, double *JVS);
  void (*solve)(const double *JVS, double *X);
} kpp_mech_t;

#endif
