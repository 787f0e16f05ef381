/*
 * pcx_abi.h -- C ABI of libpancax_b200.so: the B200 (sm_100a) implementation of pancax's
 * training hot path (deep-energy / FE-residual PINN loss + parameter gradient over every
 * quadrature point of a finite-element mesh).
 *
 * Plain C, plain pointers and sizes.  Every `const double*` / `const int32_t*` argument that
 * is documented "device" is a CUDA device pointer on the handle's device; the caller (XLA,
 * torch, ...) owns every input and output buffer, the library owns only the handle's workspace.
 * Compute calls enqueue work on `stream` (a cudaStream_t passed as void*); they never allocate,
 * never synchronise and never read results back to the host.  Outputs are fully overwritten.
 *
 * Return value: 0 on success, a negative PCX_ERR_* code otherwise; pcx_last_error() returns a
 * thread-local message.  There is no CPU fallback: an unsupported element / model / network
 * shape is an error, never a silently different code path.
 *
 * Each entry point cites the reference interface (pancax 0.0.16) it replaces.
 */
#ifndef PCX_ABI_H
#define PCX_ABI_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCX_ABI_VERSION 5

/* error codes */
#define PCX_OK               0
#define PCX_ERR_INVALID     -1  /* bad argument / unsupported shape            */
#define PCX_ERR_CUDA        -2  /* CUDA runtime error (message has the detail) */
#define PCX_ERR_STATE       -3  /* call order violated (e.g. model not set)    */
#define PCX_ERR_UNSUPPORTED -4  /* valid pancax input this library refuses (no fallback) */

/* element types: pancax/fem/read_exodus_mesh.py:28-55 */
#define PCX_ELEM_HEX8  0
#define PCX_ELEM_QUAD4 1
#define PCX_ELEM_TRI3  2

/* physics: pancax/physics_kernels/{poisson.py:8-19, solid_mechanics.py:86-109} */
#define PCX_PHYS_POISSON 0
#define PCX_PHYS_SOLID   1

/* formulations: pancax/physics_kernels/solid_mechanics.py:20-29 (PlaneStrain), :32-71 (PlaneStress), :74-83 (ThreeDimensional) */
#define PCX_FORM_PLANE_STRAIN      0
#define PCX_FORM_THREE_DIMENSIONAL 1
#define PCX_FORM_PLANE_STRESS      2   /* solid_mechanics.py:32-71: grad_u_33 = per-point root of sigma_33 = 0
                                          (math/scalar_root_find.py), differentiated implicitly */

/* constitutive models: pancax/constitutive_models/mechanics/hyperelasticity/*.py
 * property order = dataclass field order of the reference classes:
 *   NEOHOOKEAN  : bulk_modulus, shear_modulus                        (neohookean.py:17-18)
 *   GENT        : bulk_modulus, shear_modulus, Jm_parameter          (gent.py:18-20)
 *   SWANSON     : bulk_modulus, A1, P1, B1, Q1, C1, R1, cutoff_strain (swanson.py:18-26; cutoff static)
 *   HENCKY      : bulk_modulus, shear_modulus                        (hencky.py:7-8)
 *   BLATZ_KO    : shear_modulus                                      (blatz_ko.py:7)
 */
#define PCX_MODEL_NONE       0
#define PCX_MODEL_NEOHOOKEAN 1
#define PCX_MODEL_GENT       2
#define PCX_MODEL_SWANSON    3
#define PCX_MODEL_HENCKY     4
#define PCX_MODEL_BLATZ_KO   5
#define PCX_MAX_PROPS        8
#define PCX_MAX_PRONY       20

/* loss kinds: pancax/loss_functions/weak_form_loss_functions.py */
#define PCX_LOSS_ENERGY                   0  /* EnergyLoss.path_independent_call            :74-88   */
#define PCX_LOSS_ENERGY_RESIDUAL          1  /* EnergyAndResidualLoss.path_independent_call :194-205 */
#define PCX_LOSS_ENERGY_RESIDUAL_REACTION 2  /* EnergyResidualAndReactionLoss               :278-300 */

typedef struct pcx_handle pcx_handle;

/* Static inputs of the path: what VariationalDomain + ForwardProblem build
 * (pancax/domains.py:132-156, pancax/problems.py:34-96, pancax/fem/dof_manager.py:31-42,
 *  pancax/data.py:222-243). */
typedef struct {
  int32_t elem_type;          /* PCX_ELEM_*                                                   */
  int32_t q_order;            /* quadrature degree (pancax/fem/quadrature_rules.py)           */
  int32_t ndof;               /* field dofs per node (physics.n_dofs)                         */
  int32_t reaction_dof;       /* GlobalData.reaction_dof                                       */
  int64_t ne;                 /* elements                                                      */
  int64_t nn;                 /* nodes                                                         */
  int64_t n_unknown;          /* len(dof_manager.unknownIndices); 0 if residual losses unused  */
  int64_t n_reaction;         /* len(global_data.reaction_nodes); 0 if unused                  */
  const double*  coords;      /* device, f64 [nn*nd] row-major (mesh.coords)                   */
  const int32_t* conns;       /* device, i32 [ne*nnpe] row-major, 0-based (mesh.conns)          */
  const int64_t* unknown_idx; /* device, i64 [n_unknown], dof id = node*ndof+comp, or NULL (residual losses unused).
                                 A non-NULL pointer with n_unknown = 0 is the empty list: every dof constrained, R = 0
                                 with a zero gradient (what jnp.linalg.norm of an empty slice gives)            */
  const int32_t* reaction_nodes; /* device, i32 [n_reaction], or NULL; non-NULL with n_reaction = 0: reaction = 0 */
  int32_t nt;                 /* number of time steps (len(domain.times))                       */
  int32_t deterministic;      /* 1: node-centric fixed-order force assembly; 0: fp64 atomics    */
  const double* times;        /* HOST, f64 [nt]                                                 */
} pcx_mesh_desc;

/* Field = input normalisation + MLP + Dirichlet wrapper: pancax/networks/fields.py:10-101,
 * pancax/networks/mlp.py:49-85 (equinox.nn.MLP, tanh hidden activation, identity final one).
 * Weights travel as ONE flat f64 device vector in equinox leaf order
 *   [W0 (width,n_in), b0 (width), W1 (width,width), b1, ..., Wout (n_out,width) (, bout)].
 * The Dirichlet wrapper g(x,t,z) is a Python callable in pancax (fields.py:101); it must be
 * diagonal-affine in z (all shipped pancax/bvps are) and is passed tabulated at the mesh nodes:
 *   u[t][n][i] = bc_a[t][n][i] + bc_b[t][n][i] * z[n][i]
 * with bc_a = g(x,t,0), bc_b = g(x,t,e_i)_i - bc_a.  NULL tables mean g = identity. */
typedef struct {
  int32_t n_in;               /* nd + 1                                                         */
  int32_t n_out;              /* ndof                                                           */
  int32_t width;              /* n_neurons                                                      */
  int32_t depth;              /* n_layers (hidden layers)                                       */
  int32_t use_final_bias;     /* mlp.py:61 (default False)                                      */
  int32_t normalize_time;     /* fields.py:72-76                                                */
  double  x_min[3], x_max[3]; /* fields.py:70-71                                                */
  double  t_min, t_max;       /* fields.py:68-69                                                */
  const double* bc_a;         /* device, f64 [nt*nn*ndof] or NULL                               */
  const double* bc_b;         /* device, f64 [nt*nn*ndof] or NULL                               */
} pcx_field_desc;

/* Constitutive properties.  Property k is either fixed (value[k]) or a trainable
 * BoundedProperty (pancax/constitutive_models/properties.py:10-38):
 *   p = prop_min + (prop_max - prop_min) * sigmoid(raw)
 * Trainable raw values travel in a device vector `prop_raw` in property order, skipping fixed ones. */
typedef struct {
  int32_t physics;            /* PCX_PHYS_*                                                     */
  int32_t model;              /* PCX_MODEL_* (PCX_MODEL_NONE for Poisson)                       */
  int32_t formulation;        /* PCX_FORM_*                                                     */
  int32_t nprops;
  double  value[PCX_MAX_PROPS];
  int32_t bounded[PCX_MAX_PROPS];
  double  prop_min[PCX_MAX_PROPS];
  double  prop_max[PCX_MAX_PROPS];
  const double* source_q;     /* Poisson: device f64 [ne*nq], f evaluated at the quadrature
                                 points (poisson.py:17; Python callable tabulated once) or NULL */
  /* SimpleFeFv (pancax/constitutive_models/mechanics/hyperviscoelasticity/simple_fefv.py:11-14): `model` is the
   * equilibrium model; n_prony > 0 adds that many non-equilibrium Prony branches (PronySeries.moduli /
   * .relaxation_times, base.py:20-28), each with a 3x3 viscous deformation gradient per quadrature point as state.
   * The shift factor is 1, as in the reference (simple_fefv.py:27).  value[] / bounded[] describe the equilibrium
   * model's properties (trainable ones allowed).  A model with state is defined for the path-dependent calls only
   * (pcx_loss_desc.first_step = 1): EnergyLoss (weak_form_loss_functions.py:48-72), EnergyAndResidualLoss
   * (:151-192) and EnergyResidualAndReactionLoss (:242-276) -- the internal force of a step is dPi/du at fixed old
   * state (base.py:356-361), the gradient carries the second-order terms through the state history.  PlaneStress
   * with a state model (the root find then runs on the total stress at fixed old state, solid_mechanics.py:49-71) is
   * defined under the EnergyLoss; under the residual losses, in the phased (sharded) calls and in the op-level calls
   * a state model is refused. */
  int32_t n_prony;
  double  prony_moduli[PCX_MAX_PRONY];
  double  prony_times[PCX_MAX_PRONY];
} pcx_physics_desc;

typedef struct {
  int32_t kind;               /* PCX_LOSS_*                                                     */
  double  energy_weight;
  double  residual_weight;
  double  reaction_weight;
  const double* global_outputs; /* DEVICE f64 [nt] (problem.global_data.outputs) or NULL; read by the kernels
                                   of the call, never staged (ABI v4; v3 took a host pointer)       */
  int32_t first_step;         /* 0: path-independent call (vmap over all times); 1: the path-dependent call
                                 (weak_form_loss_functions.py:48-72,151-192,242-276; required for a model with state): the
                                 loop starts at time step 1 ("time step 0 is the initial condition"), residual and
                                 reaction terms are still divided by nt                                  */
} pcx_loss_desc;

/* aux layout written by pcx_loss_and_grad (device f64 [PCX_AUX_FIXED + nt]):
 *   [0] energy  [1] residual  [2] global_data_loss  [3] flags (0 = ok; bit0: det(F)<=0 or NaN seen)
 *   [4 .. 4+nt) reactions[t]
 * (keys of the reference's aux dicts: weak_form_loss_functions.py:88,205,297-300) */
#define PCX_AUX_ENERGY      0
#define PCX_AUX_RESIDUAL    1
#define PCX_AUX_GLOBAL_DATA 2
#define PCX_AUX_FLAGS       3
#define PCX_AUX_FIXED       4

int pcx_abi_version(void);
const char* pcx_last_error(void);

/* lifetime */
int pcx_create(pcx_handle** out, const pcx_mesh_desc* mesh);
int pcx_set_physics(pcx_handle* h, const pcx_physics_desc* phys);
int pcx_set_field(pcx_handle* h, const pcx_field_desc* field);
int pcx_destroy(pcx_handle* h);

/* options.  PCX_OPT_CULL_NULL_STEPS (default 1): a time step whose Dirichlet table b is identically zero
 * (every shipped load ramp at t = 0, pancax/bvps/*.py) has u = a exactly and a zero cotangent on the network
 * output, so the MLP forward / adjoint are skipped for it.  Results are identical to evaluating every row;
 * 0 evaluates every row like the reference does. */
#define PCX_OPT_CULL_NULL_STEPS 1
/* PCX_OPT_TF32_ADJOINT (default 0): run the adjoint through the MLP (delta chain + weight-gradient contraction) on
 * the TF32 tensor cores with error-compensated 3xTF32 products and fp32 accumulation flushed to fp64 every 1 024
 * rows, from fp32 copies of the hidden activations (BASELINE.json north_star item 4 / config C5).  The forward pass,
 * the element sweeps, the loss and aux stay fp64 and bit-identical; only d loss / d weights changes, by <= 1e-5
 * relative per leaf (north_star allows 1e-4).  Must not be toggled between a forward and its adjoint
 * (pcx_loss_begin / pcx_loss_finish).  Value 1: mma.sync.m16n8k8.tf32 kernel (csrc/pcx_mlp_tf32.cuh).  Value 2: every
 * contraction over rows (dW_L = d_L^T h_L, dW_out, dW_0) on tcgen05.mma kind::tf32 with TMEM accumulators
 * (csrc/pcx_mlp_umma.cuh; widths 49..55 with 2..4 hidden layers, other shapes run the value-1 kernel); same tolerance. */
#define PCX_OPT_TF32_ADJOINT 2
/* Kernel-variant switches for A/B measurements (tools/kernel_ab.py); every variant computes the same function to
 * rounding.  FWD_TANH: 1 (default) replicated conflict-free exp2 table + 13-instruction tanh, 0 single table, 2 replicated
 * table with the old range reduction.  FWD_EDGE / BWD_EDGE: the mostly-padding last feature tile of the hidden
 * layers on the FP64 vector pipe instead of the tensor pipe (BWD_EDGE: 1 = delta chain, 2 (default) = + the 13 edge blocks of
 * the weight-gradient contraction at width 50).  FWD_VARIANT: (row tiles per warp, min CTAs per SM) = 12|21|22. */
#define PCX_OPT_TUNE_FWD_TANH    100
#define PCX_OPT_TUNE_FWD_EDGE    101
#define PCX_OPT_TUNE_BWD_EDGE    102
#define PCX_OPT_TUNE_FWD_VARIANT 103
#define PCX_OPT_TUNE_FWD_DYNAMIC 105   /* 1 (default): forward row groups handed out by a global counter; 0: static grid stride */
#define PCX_OPT_TUNE_L2PF        107   /* 1 (default): the adjoint pulls its next tile's stored activations into L2 a tile ahead */
#define PCX_OPT_TUNE_STAMPS      104   /* value = device pointer to u64[3 * 2 * SM count] or 0: the MLP kernels over the mesh
                                          nodes write per-CTA (start, end, smid) %globaltimer stamps of their last launch there */
int pcx_set_option(pcx_handle* h, int32_t option, int64_t value);

/* Set-up call for workspaces whose size depends on the caller's batch (everything else is sized by pcx_set_physics /
 * pcx_set_field): PCX_RESERVE_POINTS = the strong-form / collocation path for batches of up to n points.  Allocates
 * and may synchronise: call it outside traced / captured regions.  Compute calls never allocate; one that needs more
 * than was reserved returns PCX_ERR_STATE. */
#define PCX_RESERVE_POINTS 1
int pcx_reserve(pcx_handle* h, int32_t what, int64_t n);

/* sizes the caller needs to allocate outputs */
int64_t pcx_num_weights(const pcx_handle* h);     /* flat MLP parameter count          */
int64_t pcx_num_trainable_props(const pcx_handle* h);
int64_t pcx_num_qp(const pcx_handle* h);          /* ne * nq                            */
int32_t pcx_nq(const pcx_handle* h);
int32_t pcx_nnpe(const pcx_handle* h);
int32_t pcx_nd(const pcx_handle* h);

/* (1) element gather / geometry -- NonAllocatedFunctionSpace.{shape_function_gradients,JxWs}
 *     pancax/fem/function_space.py:63-89 and x_q of base.py:322.
 *     grad_N [ne*nq*nnpe*nd], JxW [ne*nq], x_q [ne*nq*nd] (any may be NULL). */
int pcx_element_geometry(pcx_handle* h, double* grad_N, double* JxW, double* x_q, void* stream);

/* (2) field at the mesh nodes -- BasePhysics.vmap_field_values, pancax/physics_kernels/base.py:248-250.
 *     us [nn*ndof] at times[t_idx]. */
int pcx_field_forward(pcx_handle* h, const double* weights, int32_t t_idx, double* us, void* stream);

/* (4) its adjoint: g_weights [pcx_num_weights] = d<g_us, us>/d weights (VJP of (2)). */
int pcx_field_backward(pcx_handle* h, const double* weights, int32_t t_idx, const double* g_us,
                       double* g_weights, void* stream);

/* field at arbitrary points (FullFieldDataLoss, data_loss_functions.py:13-24):
 *   pts [n*(nd+1)] rows (x..., t);  u = a + b*z with optional tables a,b [n*ndof];  u [n*ndof]. */
int pcx_points_forward(pcx_handle* h, const double* weights, const double* pts, const double* a,
                       const double* b, int64_t n, double* u, void* stream);
int pcx_points_backward(pcx_handle* h, const double* weights, const double* pts, const double* b,
                        const double* g_u, int64_t n, double* g_weights, void* stream);

/* (3)+(5) potential energy and internal force -- BaseEnergyFormPhysics.potential_energy /
 *     potential_energy_and_internal_force, base.py:349-361.
 *     Pi: device scalar; f [nn*ndof] or NULL; g_props [n_trainable] = dPi/d prop_raw or NULL. */
int pcx_energy_force(pcx_handle* h, const double* prop_raw, const double* us, double* Pi,
                     double* f, double* g_props, void* stream);

/* post-processing fields at the quadrature points (what pancax's element_pp wrappers evaluate for the Exodus
 * output: pancax/physics_kernels/base.py:24-158, solid_mechanics.py:115-170):
 *   F [ne*nq*9] deformation gradient (3x3 row-major; plane strain embedded, mechanics/base.py:19-21),
 *   I1_bar [ne*nq] (mechanics/base.py:50-63), P [ne*nq*9] first Piola-Kirchhoff stress = d energy / d grad_u
 *   (mechanics/base.py:112-118).  Any output may be NULL. */
int pcx_element_fields(pcx_handle* h, const double* prop_raw, const double* us, double* F, double* I1_bar,
                       double* P, void* stream);

/* one time step of a model with state variables (SimpleFeFv) at a GIVEN old state -- the op-level calls of the reference
 * for such a model: physics.potential_energy(params, domain, t, us, state_old, dt) -> (Pi, state_new)
 * (pancax/physics_kernels/base.py:349-354), potential_energy_and_internal_force(..., state_old, dt) (f = dPi/du at fixed
 * old state, base.py:356-361) and the post-processor's element_pp(pk1_stress)(..., us, theta, state_old, dt)
 * (base.py:24-74, constitutive_models/mechanics/base.py:112-118; driven step by step in post_processor.py:244-354,470-490).
 *   state_old, state_new [ne*nq*n_prony*9] (must not alias), dt >= 0 (dt = 0, the post-processor's step 0: the reference
 *   evaluates 0/0 there; the limit dt -> 0 is returned), Pi device scalar or NULL, f [nn*ndof] or NULL,
 *   P [ne*nq*9] (3x3 row-major) or NULL.  PlaneStress (solid_mechanics.py:49-71): the out-of-plane stretch of every
 *   point is the root of the total sigma_33 at the given old state; f and P are evaluated at those roots. */
int pcx_state_step(pcx_handle* h, const double* prop_raw, const double* us, const double* state_old, double dt,
                   double* state_new, double* Pi, double* f, double* P, void* stream);

/* tangent-stiffness action Kv = d/d(eps) f(us + eps v): the second-order term the reference gets
 * by differentiating through value_and_grad in base.py:363-385.  g_props = d<f,v>/d prop_raw or NULL. */
int pcx_stiffness_action(pcx_handle* h, const double* prop_raw, const double* us, const double* v,
                         double* Kv, double* g_props, void* stream);

/* residual norm and reaction force of an assembled f -- base.py:369-371, :381-384.
 *   out[0] = ||f[unknown]||_2, out[1] = sum f[reaction_nodes, reaction_dof]. */
int pcx_residual_reaction(pcx_handle* h, const double* f, double* out, void* stream);

/* fused fast path: loss, aux and d loss / d (weights, prop_raw) for the built-in weak-form losses --
 * eqx.filter_value_and_grad(loss.filtered_loss), pancax/optimizers.py:77-88.
 *   loss: device scalar; aux: device [PCX_AUX_FIXED+nt]; g_weights [num_weights] or NULL (value and aux
 *   only: what calling the loss without grad does); g_props [n_trainable] or NULL. */
int pcx_loss_and_grad(pcx_handle* h, const pcx_loss_desc* loss, const double* weights,
                      const double* prop_raw, double* loss_out, double* aux, double* g_weights,
                      double* g_props, void* stream);

/* Sharded residual / reaction losses (one shard of the elements per GPU): pcx_loss_and_grad cut at the two
 * points where shards must talk to each other (reference semantics are still
 * weak_form_loss_functions.py:194-205,278-300 on the WHOLE mesh).
 *   pcx_set_ownership: entries [0, n_unknown_owned) of unknown_idx and [0, n_reaction_owned) of
 *     reaction_nodes belong to this shard (they enter the norm / reaction sums exactly once over all
 *     shards); the remaining entries are closure copies owned elsewhere (they still carry v).
 *   pcx_loss_begin : field forward + energy/force sweep over the shard's elements; f_partial
 *     [nt*nn*ndof] (caller's device buffer) = the shard's partial internal force.
 *   -- caller: f_full = f_partial with the rows of shared nodes summed over shards (may be in place).
 *   pcx_loss_mid   : rr_partial [nt*2] (device) = (sum of f_full^2 over OWNED unknown dofs, sum of f_full
 *     over OWNED reaction nodes) per time step.
 *   -- caller: rr_total = sum over shards of rr_partial (may be in place).
 *   pcx_loss_finish: v, tangent action, adjoint.  Outputs are the shard's PARTIAL loss / aux / gradients:
 *     their SUM over shards is the value pcx_loss_and_grad returns on the whole mesh, provided
 *     replicated_scale = 1 / n_shards (it weights the terms all shards compute identically).
 *     Exception: aux[PCX_AUX_FLAGS] is the shard's own flag word (a bit mask): combine it with OR / MAX, or read a
 *     summed value as "non-zero = anomaly on some shard". */
int pcx_set_ownership(pcx_handle* h, int64_t n_unknown_owned, int64_t n_reaction_owned);
int pcx_loss_begin(pcx_handle* h, const pcx_loss_desc* loss, const double* weights, const double* prop_raw,
                   double* f_partial, void* stream);
int pcx_loss_mid(pcx_handle* h, const double* f_full, double* rr_partial, void* stream);
int pcx_loss_finish(pcx_handle* h, const pcx_loss_desc* loss, const double* f_full, const double* rr_total,
                    double replicated_scale, double* loss_out, double* aux, double* g_weights, double* g_props,
                    void* stream);

/* Pack / unpack kernels of the exchange between pcx_loss_begin and pcx_loss_mid: rows idx[0..n) (distinct node ids) of
 * f [nt][nn][ndof] <-> a contiguous buffer [nt][n][ndof].  pcx_scatter_rows: accumulate = 0 overwrites (in = NULL: zeros),
 * accumulate = 1 adds.  The caller sends the packed rows to the shards that share those nodes (ncclSend/Recv between
 * neighbours) and adds what it receives in ascending shard order, so every sharer ends with bit-identical sums. */
int pcx_gather_rows(const double* f, int64_t nt, int64_t nn, int32_t ndof, const int64_t* idx, int64_t n, double* out,
                    void* stream);
int pcx_scatter_rows(double* f, int64_t nt, int64_t nn, int32_t ndof, const int64_t* idx, int64_t n, const double* in,
                     int32_t accumulate, void* stream);

/* FullFieldDataLoss value + weight gradient on one batch -- data_loss_functions.py:13-24.
 *   loss: device scalar = weight * mean((u - outputs)^2). */
int pcx_field_data_loss_and_grad(pcx_handle* h, const double* weights, const double* pts,
                                 const double* a, const double* b, const double* outputs, int64_t n,
                                 double weight, double* loss_out, double* g_weights, void* stream);

/* Strong-form / collocation path: the jet of the Field at arbitrary points -- what pancax derives
 * with nested jacfwd / hessian (pancax/physics_kernels/base.py:205-229: field_values, field_gradients,
 * field_laplacians = trace of field_hessians, field_time_derivatives) -- evaluated in ONE fused pass per point
 * (forward-mode coordinate tangents through the MLP on the FP64 tensor pipe).
 *   pts [n*(nd+1)] rows (x..., t); tab [n*ndof*2*(nd+3)] or NULL: the Dirichlet wrapper u = a + b z (fields.py:101)
 *   and its derivatives at the points, per (point, dof): (a, d_j a.., lap a, a_t, b, d_j b.., lap b, b_t); NULL = identity.
 *   Outputs (any may be NULL): u [n*ndof], grad_u [n*ndof*nd], lap_u [n*ndof], u_t [n*ndof].  The d/dt stream (one of
 *   nd + 3) is carried only when u_t is asked for; the other outputs are bit-identical either way. */
int pcx_points_jet(pcx_handle* h, const double* weights, const double* pts, const double* tab, int64_t n,
                   double* u, double* grad_u, double* lap_u, double* u_t, void* stream);

/* Strong-form residual loss and its weight gradient on a batch of collocation points:
 *   r = c_t u_t - c_lap lap(u) - src - target,   loss = weight * mean over (points, dofs) of r^2
 * Poisson (poisson.py:26-29): c_t = 0, c_lap = 1, src = f(x); heat (heat_equation.py:11-17): c_t = rho c, c_lap = k;
 * StrongFormResidualLoss (strong_form_loss_functions.py:8-27) when every time step carries the same points;
 * `target` is the collocation example's `outputs` (NULL = 0).  src / target [n*ndof] or NULL.
 *   loss_out: device scalar; resid_out [n*ndof] (r before the target is subtracted) or NULL; g_weights
 *   [pcx_num_weights] or NULL (value only).  c_t == 0 (no time term in the residual): the d/dt stream is not carried. */
int pcx_strong_form_loss_and_grad(pcx_handle* h, const double* weights, const double* pts, const double* tab,
                                  const double* src, const double* target, int64_t n, double c_t, double c_lap,
                                  double weight, double* loss_out, double* resid_out, double* g_weights, void* stream);

/* State variables after a path-dependent call of a model with state (SimpleFeFv): the viscous deformation gradients
 * Fv of time step t_idx, device f64 [ne*nq*n_prony*9] in the reference's state layout (ne, nq, n_prony*9) --
 * what pancax's post-processor writes as `state_variables` (solid_mechanics.py:157-168, is_state_method).  Valid
 * after pcx_loss_and_grad (first_step = 1) on this handle; step 0 is initial_state() (identities). */
int64_t pcx_num_state_variables(const pcx_handle* h);   /* n_prony * 9 per quadrature point (0: no state) */
int pcx_get_state(pcx_handle* h, int32_t t_idx, double* state, void* stream);

/* number of kernels this library has launched on the handle so far (for bench.py "gpu_launches") */
int64_t pcx_launch_count(const pcx_handle* h);

/* Adam update of one flat leaf on the device -- optax.chain(scale_by_adam(), scale_by_schedule(..),
 * scale(-1)) as composed in pancax/optimizers.py:155-173.  `count` is the 1-based update number (bias
 * correction), `lr` the schedule value for this update.  params/mu/nu are updated in place. */
int pcx_adam_step(double* params, const double* grads, double* mu, double* nu, int64_t n, int64_t count,
                  double lr, double b1, double b2, double eps, void* stream);

/* The same update with the update counter and the exponential-decay schedule evaluated on the device
 * (lr_k = lr0 * decay_rate^(k / transition_steps), optax.exponential_decay as used in
 * pancax/optimizers.py:155-163): no argument depends on host state, so loss + gradient + update can be
 * captured once in a CUDA graph and replayed.  state: device f64[1] = updates done so far; `advance` != 0 on
 * the last leaf of a step increments it. */
int pcx_adam_step_sched(double* params, const double* grads, double* mu, double* nu, int64_t n, double* state,
                        int32_t advance, double lr0, double decay_rate, double transition_steps, double b1,
                        double b2, double eps, void* stream);

/* measurement helpers (no reference counterpart): per-kernel-class device time from CUDA events
 * recorded on the launching stream, and the live FP64 tensor-pipe (DMMA) peak of the device. */
#define PCX_PROF_PACK            0
#define PCX_PROF_MLP_FORWARD     1
#define PCX_PROF_ELEMENT_FORCE   2
#define PCX_PROF_ASSEMBLE        3
#define PCX_PROF_ELEMENT_TANGENT 4
#define PCX_PROF_MLP_BACKWARD    5
#define PCX_PROF_REDUCE_WGRAD    6
#define PCX_PROF_NCLASS          8
int pcx_profile_enable(pcx_handle* h, int on);
int pcx_profile_collect(pcx_handle* h, double* ms_by_class, int64_t* launches_by_class);
int pcx_measure_dmma_peak(double* tflops, int reps);

#ifdef __cplusplus
}
#endif
#endif /* PCX_ABI_H */
