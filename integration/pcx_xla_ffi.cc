// pcx_xla_ffi.cc -- XLA custom-call (jax.ffi) shim over the C ABI of libpancax_b200.so (include/pcx_abi.h).
//
// BASELINE.json north_star: the pancax Python API keeps "calling hand-written sm_100a CUDA kernels through a thin C-ABI
// layer exposed as XLA custom calls via jax.ffi".  Every handler below decodes XLA buffers / attributes and forwards to
// exactly one ABI entry point on the stream XLA hands it; no handler allocates, synchronises or touches the host data of
// a buffer (the ABI's own contract), so the calls are safe inside jit, scan and XLA's command buffers (CUDA graphs).
//
// Build (on a box with jaxlib; NOT possible in this repository's build image -- no jax wheels, no network):
//   g++ -O2 -std=c++17 -fPIC -shared integration/pcx_xla_ffi.cc -o libpcx_xla_ffi.so \
//       -I$(python -c "import jax.ffi as f; print(f.include_dir())") -Iinclude -I/usr/local/cuda/include \
//       -Lpancax_b200/lib -lpancax_b200 -Wl,-rpath,'$ORIGIN'
// Here it is type-checked against a declaration-only stand-in for the XLA header (integration/stub/xla/ffi/api/ffi.h):
//   g++ -std=c++17 -fsyntax-only -Iintegration/stub -Iinclude -I/usr/local/cuda/include integration/pcx_xla_ffi.cc
// (tests/test_integration_sources.py).  Python side: integration/kernels_b200.py.
//
// Conventions: the pcx_handle travels as an int64 attribute "handle" (its address; created / configured on the host by
// kernels_b200.py through ctypes: pcx_create / pcx_set_physics / pcx_set_field / pcx_reserve are set-up calls, not
// custom calls).  Optional inputs are passed as zero-length buffers and forwarded as NULL.  Flat weights are the leaves
// of params.fields.networks in equinox order (pcx_field_desc).  Non-zero return codes become ffi::Error::Internal with
// pcx_last_error() as the message: there is no fallback path.
#include <cuda_runtime_api.h>

#include <cstdint>
#include <string>

#include "pcx_abi.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;
using F64 = ffi::Buffer<ffi::F64>;
using S32 = ffi::Buffer<ffi::S32>;
using S64 = ffi::Buffer<ffi::S64>;
using RF64 = ffi::ResultBuffer<ffi::F64>;

namespace {

inline ffi::Error pcx_err(int rc) {
  return rc == PCX_OK ? ffi::Error::Success() : ffi::Error::Internal(std::string("libpancax_b200: ") + pcx_last_error());
}
inline pcx_handle* H(int64_t h) { return reinterpret_cast<pcx_handle*>(static_cast<intptr_t>(h)); }
inline const double* opt(const F64& b) { return b.element_count() ? b.typed_data() : nullptr; }   // zero-length = absent
inline double* opt(RF64& b) { return b->element_count() ? b->typed_data() : nullptr; }
inline ffi::Error copy_if_distinct(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  // in-place ABI calls (Adam) when XLA did not alias the output to the input (input_output_aliases missing)
  if (dst == src || bytes == 0) return ffi::Error::Success();
  return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s) == cudaSuccess
             ? ffi::Error::Success() : ffi::Error::Internal("cudaMemcpyAsync failed");
}

// ---------------------------------------------------------------------------------------------- (1) geometry
// NonAllocatedFunctionSpace.shape_function_gradients / JxWs, x_q (pancax/fem/function_space.py:63-89)
ffi::Error ElementGeometryImpl(cudaStream_t s, int64_t h, RF64 grad_N, RF64 JxW, RF64 x_q) {
  return pcx_err(pcx_element_geometry(H(h), opt(grad_N), opt(JxW), opt(x_q), s));
}

// ---------------------------------------------------------------------------------------------- (2)/(4) field
// BasePhysics.vmap_field_values (pancax/physics_kernels/base.py:248-250) and its VJP
ffi::Error FieldForwardImpl(cudaStream_t s, int64_t h, int32_t t_idx, F64 weights, RF64 us) {
  return pcx_err(pcx_field_forward(H(h), weights.typed_data(), t_idx, us->typed_data(), s));
}
ffi::Error FieldBackwardImpl(cudaStream_t s, int64_t h, int32_t t_idx, F64 weights, F64 g_us, RF64 g_w) {
  return pcx_err(pcx_field_backward(H(h), weights.typed_data(), t_idx, g_us.typed_data(), g_w->typed_data(), s));
}
// field at arbitrary (x, t) rows (FullFieldDataLoss, data_loss_functions.py:13-24); a / b zero-length = identity wrapper
ffi::Error PointsForwardImpl(cudaStream_t s, int64_t h, F64 weights, F64 pts, F64 a, F64 b, RF64 u) {
  const int64_t n = static_cast<int64_t>(pts.dimensions()[0]);
  return pcx_err(pcx_points_forward(H(h), weights.typed_data(), pts.typed_data(), opt(a), opt(b), n, u->typed_data(), s));
}
ffi::Error PointsBackwardImpl(cudaStream_t s, int64_t h, F64 weights, F64 pts, F64 b, F64 g_u, RF64 g_w) {
  const int64_t n = static_cast<int64_t>(pts.dimensions()[0]);
  return pcx_err(pcx_points_backward(H(h), weights.typed_data(), pts.typed_data(), opt(b), g_u.typed_data(), n,
                                     g_w->typed_data(), s));
}

// ---------------------------------------------------------------------------------------------- (3)/(5) energy, force
// BaseEnergyFormPhysics.potential_energy[_and_internal_force] (base.py:349-361); f zero-length = energy only
ffi::Error EnergyForceImpl(cudaStream_t s, int64_t h, F64 prop_raw, F64 us, RF64 Pi, RF64 f, RF64 g_props) {
  return pcx_err(pcx_energy_force(H(h), opt(prop_raw), us.typed_data(), Pi->typed_data(), opt(f), opt(g_props), s));
}
// K v: the second-order term of base.py:363-385 (symmetric for a hyperelastic W: also the VJP of f)
ffi::Error StiffnessActionImpl(cudaStream_t s, int64_t h, F64 prop_raw, F64 us, F64 v, RF64 Kv, RF64 g_props) {
  return pcx_err(pcx_stiffness_action(H(h), opt(prop_raw), us.typed_data(), v.typed_data(), Kv->typed_data(), opt(g_props), s));
}
// (||f[unknown]||, sum f[reaction_nodes, reaction_dof]) (base.py:369-371, :381-384)
ffi::Error ResidualReactionImpl(cudaStream_t s, int64_t h, F64 f, RF64 out) {
  return pcx_err(pcx_residual_reaction(H(h), f.typed_data(), out->typed_data(), s));
}
// post-processor element variables F, I1_bar, P (physics_kernels/base.py:24-158, solid_mechanics.py:115-170)
ffi::Error ElementFieldsImpl(cudaStream_t s, int64_t h, F64 prop_raw, F64 us, RF64 F, RF64 I1_bar, RF64 P) {
  return pcx_err(pcx_element_fields(H(h), opt(prop_raw), us.typed_data(), opt(F), opt(I1_bar), opt(P), s));
}

// ---------------------------------------------------------------------------------------------- fused loss + gradient
// eqx.filter_value_and_grad(loss.filtered_loss, has_aux=True) (pancax/optimizers.py:77-88) for the built-in weak-form
// losses (weak_form_loss_functions.py:74-88,194-205,278-300; first_step = 1: the path-dependent calls :48-72,242-276).
// global_outputs: device f64[nt] (problem.global_data.outputs) or zero-length.  want_grad = 0: value and aux only.
ffi::Error LossAndGradImpl(cudaStream_t s, int64_t h, int32_t kind, int32_t first_step, int32_t want_grad, double energy_weight,
                           double residual_weight, double reaction_weight, F64 weights, F64 prop_raw, F64 global_outputs,
                           RF64 loss, RF64 aux, RF64 g_w, RF64 g_props) {
  pcx_loss_desc d;
  d.kind = kind;
  d.energy_weight = energy_weight;
  d.residual_weight = residual_weight;
  d.reaction_weight = reaction_weight;
  d.global_outputs = opt(global_outputs);
  d.first_step = first_step;
  return pcx_err(pcx_loss_and_grad(H(h), &d, weights.typed_data(), opt(prop_raw), loss->typed_data(), aux->typed_data(),
                                   want_grad ? g_w->typed_data() : nullptr, opt(g_props), s));
}

// Sharded residual / reaction losses: pcx_loss_and_grad cut at the two exchange points; kernels_b200.py puts a
// jax.lax.psum over the device axis of a shard_map between the three calls.
ffi::Error LossBeginImpl(cudaStream_t s, int64_t h, int32_t kind, double energy_weight, double residual_weight,
                         double reaction_weight, F64 weights, F64 prop_raw, F64 global_outputs, RF64 f_partial) {
  pcx_loss_desc d{kind, energy_weight, residual_weight, reaction_weight, opt(global_outputs), 0};
  return pcx_err(pcx_loss_begin(H(h), &d, weights.typed_data(), opt(prop_raw), f_partial->typed_data(), s));
}
ffi::Error LossMidImpl(cudaStream_t s, int64_t h, F64 f_full, RF64 rr_partial) {
  return pcx_err(pcx_loss_mid(H(h), f_full.typed_data(), rr_partial->typed_data(), s));
}
ffi::Error LossFinishImpl(cudaStream_t s, int64_t h, int32_t kind, double energy_weight, double residual_weight,
                          double reaction_weight, double replicated_scale, F64 global_outputs, F64 f_full, F64 rr_total,
                          RF64 loss, RF64 aux, RF64 g_w, RF64 g_props) {
  pcx_loss_desc d{kind, energy_weight, residual_weight, reaction_weight, opt(global_outputs), 0};
  return pcx_err(pcx_loss_finish(H(h), &d, f_full.typed_data(), rr_total.typed_data(), replicated_scale, loss->typed_data(),
                                 aux->typed_data(), g_w->typed_data(), opt(g_props), s));
}

// FullFieldDataLoss value + weight gradient on one batch (data_loss_functions.py:13-24)
ffi::Error FieldDataLossAndGradImpl(cudaStream_t s, int64_t h, double weight, F64 weights, F64 pts, F64 a, F64 b, F64 outputs,
                                    RF64 loss, RF64 g_w) {
  const int64_t n = static_cast<int64_t>(pts.dimensions()[0]);
  return pcx_err(pcx_field_data_loss_and_grad(H(h), weights.typed_data(), pts.typed_data(), opt(a), opt(b), outputs.typed_data(),
                                              n, weight, loss->typed_data(), g_w->typed_data(), s));
}

// ---------------------------------------------------------------------------------------------- strong form
// field_values / field_gradients / field_laplacians / field_time_derivatives (physics_kernels/base.py:205-229)
ffi::Error PointsJetImpl(cudaStream_t s, int64_t h, F64 weights, F64 pts, F64 tab, RF64 u, RF64 grad_u, RF64 lap_u, RF64 u_t) {
  const int64_t n = static_cast<int64_t>(pts.dimensions()[0]);
  return pcx_err(pcx_points_jet(H(h), weights.typed_data(), pts.typed_data(), opt(tab), n, opt(u), opt(grad_u), opt(lap_u),
                                opt(u_t), s));
}
// StrongFormResidualLoss (strong_form_loss_functions.py:8-27) with the Poisson / heat residual
ffi::Error StrongFormLossAndGradImpl(cudaStream_t s, int64_t h, double c_t, double c_lap, double weight, int32_t want_grad,
                                     F64 weights, F64 pts, F64 tab, F64 src, F64 target, RF64 loss, RF64 resid, RF64 g_w) {
  const int64_t n = static_cast<int64_t>(pts.dimensions()[0]);
  return pcx_err(pcx_strong_form_loss_and_grad(H(h), weights.typed_data(), pts.typed_data(), opt(tab), opt(src), opt(target), n,
                                               c_t, c_lap, weight, loss->typed_data(), opt(resid),
                                               want_grad ? g_w->typed_data() : nullptr, s));
}

// ---------------------------------------------------------------------------------------------- state, optimiser
// state_variables element output after a path-dependent call (solid_mechanics.py:157-168)
ffi::Error GetStateImpl(cudaStream_t s, int64_t h, int32_t t_idx, RF64 state) {
  return pcx_err(pcx_get_state(H(h), t_idx, state->typed_data(), s));
}
// one time step of a state-variable model at a given old state: physics.potential_energy(params, domain, t, us, state_old, dt)
// -> (Pi, state_new), the internal force at fixed old state and element_pp(pk1_stress) (physics_kernels/base.py:24-74,349-361)
ffi::Error StateStepImpl(cudaStream_t s, int64_t h, double dt, F64 prop_raw, F64 us, F64 state_old, RF64 state_new, RF64 Pi,
                         RF64 f, RF64 P) {
  return pcx_err(pcx_state_step(H(h), opt(prop_raw), us.typed_data(), state_old.typed_data(), dt, state_new->typed_data(),
                                opt(Pi), opt(f), opt(P), s));
}
// optax.chain(scale_by_adam, scale_by_schedule, scale(-1)) + apply_updates (pancax/optimizers.py:142-180).  params / mu /
// nu are updated in place by the ABI: declare input_output_aliases={0: 0, 2: 1, 3: 2} in ffi_call; without aliasing the
// inputs are first copied into the outputs.
ffi::Error AdamStepImpl(cudaStream_t s, int64_t count, double lr, double b1, double b2, double eps, F64 params, F64 grads,
                        F64 mu, F64 nu, RF64 params_out, RF64 mu_out, RF64 nu_out) {
  const size_t bytes = params.size_bytes();
  for (auto pr : {std::pair<void*, const void*>{params_out->untyped_data(), params.untyped_data()},
                  std::pair<void*, const void*>{mu_out->untyped_data(), mu.untyped_data()},
                  std::pair<void*, const void*>{nu_out->untyped_data(), nu.untyped_data()}}) {
    ffi::Error e = copy_if_distinct(pr.first, pr.second, bytes, s);
    if (e.failure()) return e;
  }
  return pcx_err(pcx_adam_step(params_out->typed_data(), grads.typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                               static_cast<int64_t>(params.element_count()), count, lr, b1, b2, eps, s));
}
// the same with the update counter and optax.exponential_decay on the device (state: f64[1], advanced when `advance`)
ffi::Error AdamStepSchedImpl(cudaStream_t s, int32_t advance, double lr0, double decay_rate, double transition_steps, double b1,
                             double b2, double eps, F64 params, F64 grads, F64 mu, F64 nu, F64 state, RF64 params_out,
                             RF64 mu_out, RF64 nu_out, RF64 state_out) {
  const size_t bytes = params.size_bytes();
  for (auto pr : {std::pair<void*, const void*>{params_out->untyped_data(), params.untyped_data()},
                  std::pair<void*, const void*>{mu_out->untyped_data(), mu.untyped_data()},
                  std::pair<void*, const void*>{nu_out->untyped_data(), nu.untyped_data()}}) {
    ffi::Error e = copy_if_distinct(pr.first, pr.second, bytes, s);
    if (e.failure()) return e;
  }
  ffi::Error e = copy_if_distinct(state_out->untyped_data(), state.untyped_data(), sizeof(double), s);
  if (e.failure()) return e;
  return pcx_err(pcx_adam_step_sched(params_out->typed_data(), grads.typed_data(), mu_out->typed_data(), nu_out->typed_data(),
                                     static_cast<int64_t>(params.element_count()), state_out->typed_data(), advance, lr0,
                                     decay_rate, transition_steps, b1, b2, eps, s));
}

using Stream = ffi::PlatformStream<cudaStream_t>;

}  // namespace

#define PCX_BIND() ffi::Ffi::Bind().Ctx<Stream>().Attr<int64_t>("handle")

XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxElementGeometry, ElementGeometryImpl,
                              PCX_BIND().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxFieldForward, FieldForwardImpl,
                              PCX_BIND().Attr<int32_t>("t_idx").Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxFieldBackward, FieldBackwardImpl,
                              PCX_BIND().Attr<int32_t>("t_idx").Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxPointsForward, PointsForwardImpl,
                              PCX_BIND().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxPointsBackward, PointsBackwardImpl,
                              PCX_BIND().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxEnergyForce, EnergyForceImpl,
                              PCX_BIND().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxStiffnessAction, StiffnessActionImpl,
                              PCX_BIND().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxResidualReaction, ResidualReactionImpl, PCX_BIND().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxElementFields, ElementFieldsImpl,
                              PCX_BIND().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxLossAndGrad, LossAndGradImpl,
                              PCX_BIND().Attr<int32_t>("kind").Attr<int32_t>("first_step").Attr<int32_t>("want_grad")
                                  .Attr<double>("energy_weight").Attr<double>("residual_weight").Attr<double>("reaction_weight")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxLossBegin, LossBeginImpl,
                              PCX_BIND().Attr<int32_t>("kind").Attr<double>("energy_weight").Attr<double>("residual_weight")
                                  .Attr<double>("reaction_weight").Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxLossMid, LossMidImpl, PCX_BIND().Arg<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxLossFinish, LossFinishImpl,
                              PCX_BIND().Attr<int32_t>("kind").Attr<double>("energy_weight").Attr<double>("residual_weight")
                                  .Attr<double>("reaction_weight").Attr<double>("replicated_scale")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxFieldDataLossAndGrad, FieldDataLossAndGradImpl,
                              PCX_BIND().Attr<double>("weight").Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxPointsJet, PointsJetImpl,
                              PCX_BIND().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxStrongFormLossAndGrad, StrongFormLossAndGradImpl,
                              PCX_BIND().Attr<double>("c_t").Attr<double>("c_lap").Attr<double>("weight").Attr<int32_t>("want_grad")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxGetState, GetStateImpl, PCX_BIND().Attr<int32_t>("t_idx").Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxStateStep, StateStepImpl,
                              PCX_BIND().Attr<double>("dt").Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxAdamStep, AdamStepImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Attr<int64_t>("count").Attr<double>("lr").Attr<double>("b1")
                                  .Attr<double>("b2").Attr<double>("eps")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
XLA_FFI_DEFINE_HANDLER_SYMBOL(PcxAdamStepSched, AdamStepSchedImpl,
                              ffi::Ffi::Bind().Ctx<Stream>().Attr<int32_t>("advance").Attr<double>("lr0").Attr<double>("decay_rate")
                                  .Attr<double>("transition_steps").Attr<double>("b1").Attr<double>("b2").Attr<double>("eps")
                                  .Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>().Arg<F64>()
                                  .Ret<F64>().Ret<F64>().Ret<F64>().Ret<F64>());
