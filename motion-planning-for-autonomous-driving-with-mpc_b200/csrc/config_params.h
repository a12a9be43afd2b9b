// config_params.h -- mpcb200_config (C ABI) -> ParamsT<T> (kernel constants), plus the defaults.
#pragma once
#include "../../include/mpcb200.h"
#include "mpc_types.cuh"

namespace mpcb200 {

template <typename T>
inline ParamsT<T> params_from_config(const mpcb200_config& c) {
  ParamsT<T> p;
  p.N = c.N; p.max_iter = c.max_iter; p.hessian = c.hessian; p.ls_max = c.ls_max;
  p.dt = (T)c.dt; p.l_wb = (T)c.l_wb; p.l_fric = (T)c.l_fric;
  for (int i = 0; i < 5; ++i) p.Q[i] = (T)c.Q[i];
  for (int i = 0; i < 2; ++i) p.R[i] = (T)c.R[i];
  p.dd_min = (T)c.deltav_min; p.dd_max = (T)c.deltav_max; p.a_max = (T)c.a_max;
  p.de_min = (T)c.delta_min; p.de_max = (T)c.delta_max; p.v_min = (T)c.v_min; p.v_max = (T)c.v_max;
  p.r_sum = (T)c.r_sum; p.ego_off = (T)c.ego_offset;
  p.mu0 = (T)c.mu0; p.mu_min = (T)c.mu_min; p.mu_factor = (T)c.mu_factor;
  p.tol_step = (T)c.tol_step; p.tol_feas = (T)c.tol_feas; p.tau_min = (T)c.tau_min; p.bound_push = (T)c.bound_push;
  p.acc_factor = (T)c.acc_factor; p.acc_iters = c.acc_iters; p.stall_iters = c.stall_iters; p.trust_step = (T)c.trust_step; p.screen_inv_curv = (T)c.screen_inv_curv; p.init_rollout = c.init_rollout; p.kappa_sigma = (T)c.kappa_sigma; p.mu_min_alpha = (T)c.mu_min_alpha;
  p.stiff_slack = (T)c.stiff_slack;
  p.mu_warm = (T)c.mu_warm; p.warm_push = (T)c.warm_push; p.kappa_warm = (T)c.kappa_warm;
  p.mu_up_alpha = (T)c.mu_up_alpha; p.mu_up_factor = (T)c.mu_up_factor; p.mu_max = (T)c.mu_max; p.mu_factor_full = (T)c.mu_factor_full;
  return p;
}

inline void default_config(mpcb200_config* c, int N, int precision) {
  c->abi_version = MPCB200_ABI_VERSION; c->device = 0; c->N = N; c->max_batch = 4096;
  c->precision = precision; c->hessian = MPCB200_HESS_GAUSS_NEWTON; c->max_iter = 100; c->ls_max = 8;
  c->dt = 0.1; c->l_wb = 2.5789128; c->l_fric = 2.578;
  const double Q[5] = {2.3, 2.3, 500.0, 0.1, 10.0}, R[2] = {2.0, 0.2};   // config_LF_ZAM_Over-1_1.yaml:19-31
  for (int i = 0; i < 5; ++i) c->Q[i] = Q[i];
  for (int i = 0; i < 2; ++i) c->R[i] = R[i];
  c->deltav_min = -0.4; c->deltav_max = 0.4; c->a_max = 11.5; c->delta_min = -1.066; c->delta_max = 1.066;
  c->v_min = 0.0; c->v_max = 50.8;
  c->r_sum = 1.2; c->ego_offset = 0.75;                                    // lane following: dummy obstacle (Q11)
  const double ob[6] = {-100.0, 0.0, -100.0, 0.0, -100.0, 0.0};
  for (int i = 0; i < 6; ++i) c->obstacle[i] = ob[i];
  c->mu0 = 0.01; c->mu_factor = 0.2; c->tau_min = 0.99; c->bound_push = 1e-2;
  if (precision == MPCB200_F64) { c->mu_min = 1e-9; c->tol_step = 1e-8; c->tol_feas = 1e-8; c->acc_factor = 1000.0; }
  else                          { c->mu_min = 1e-7; c->tol_step = 2e-5; c->tol_feas = 1e-4; c->acc_factor = 5.0; }   // mu_min: a weakly active row ends s ~ sqrt(mu_min / h) inside its bound (5e-4 at h = 0.4; 1.6e-3 at 1e-6)
  c->acc_iters = 4; c->stall_iters = 10; c->refine_f64 = (precision == MPCB200_F32) ? 1 : 0; c->trust_step = 1e-2; c->screen_inv_curv = 1e6; c->init_rollout = 0; c->kappa_sigma = 1e10; c->mu_min_alpha = 0.5;
  c->mu_up_alpha = 0.5; c->mu_up_factor = 10.0; c->mu_max = 1e3; c->mu_factor_full = 0.04;
  c->mu_warm = 1e-4; c->warm_push = 1e-6; c->kappa_warm = 1e2; c->warm_duals = 0;
  c->warps_per_cta = 0; c->host_route = 0; c->host_chunks = 0; c->stiff_slack = 1e-3;
}

}  // namespace mpcb200
