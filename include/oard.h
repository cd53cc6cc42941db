/* oard.h — C ABI of the B200-native OA-ReactDiff denoising hot path (liboard_b200.so).
 *
 * The library replaces, for this path only, what the reference delegates to eager PyTorch + torch_scatter + PyG:
 *   oa_reactdiff/model/leftnet.py:724-891   LEFTNet.forward            -> oard_forward
 *   oa_reactdiff/model/leftnet.py:594-688   LEFTNet.__init__ / weights  -> oard_create, oard_set_weight, oard_commit_weights
 *   oa_reactdiff/utils/_graph_tools.py:9-36 edge list of the batch      -> oard_plan (consumes that edge list)
 * The reference has no FFI of its own (it is pure Python); the seam a maintainer binds is the `model=` plugin
 * argument of EGNNDynamics (oa_reactdiff/dynamics/_base.py:21,62-64).  INTEGRATION.md shows the ctypes stub.
 *
 * Conventions: plain C, no exceptions across the boundary; every entry point returns 0 on success or a negative
 * OARD_E* code and records a message retrievable with oard_last_error().  All tensor pointers passed to
 * oard_forward are DEVICE pointers owned by the caller, fp32 row-major, valid until the stream reaches the end of the
 * call's work; the call is asynchronous on `stream` (a cudaStream_t passed as void*, blocking or not).  Entry points run on
 * the handle's device and restore the caller's current device before returning.  One handle per device; a handle
 * is not thread-safe.  There is no CPU fallback: without a CUDA device every compute entry point fails with
 * OARD_ECUDA.
 */
#ifndef OARD_H
#define OARD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OARD_OK 0
#define OARD_EINVAL (-1)      /* bad argument / unsupported configuration */
#define OARD_ECUDA (-2)       /* CUDA runtime error (message holds cudaGetErrorString) */
#define OARD_ESTATE (-3)      /* call order violated (e.g. forward before plan / commit) */
#define OARD_EGRAPH (-4)      /* edge list is not CSR-sorted / not symmetric / component too large */
#define OARD_EMISSING (-5)    /* a required weight was not provided */

typedef struct oard_handle oard_handle;

/* Mirror of LEFTNet.__init__ kwargs (leftnet.py:594-611).  Only legacy=1, pos_grad=0, single_layer_output=1,
 * for_conf=0, ff=0 are implemented (the trained configuration, trainer/train_ts1x.py:43-56). */
typedef struct oard_cfg {
  int32_t hidden_channels;    /* H, <= 256 */
  int32_t num_radial;         /* R */
  int32_t num_layers;         /* L */
  int32_t in_hidden_channels; /* C, <= 32 */
  float cutoff;
  int32_t reflect_equiv;
  int32_t legacy;
  int32_t update;
  int32_t object_aware;
} oard_cfg;

int oard_abi_version(void);
const char* oard_last_error(void);

int oard_create(const oard_cfg* cfg, int device, oard_handle** out);
void oard_destroy(oard_handle* h);

/* Number of weights the configuration requires and their reference state-dict names,
 * e.g. "gcl_layers.0.edge_mlp.mlp.0.linear.weight". */
int oard_num_weights(const oard_handle* h);
const char* oard_weight_name(const oard_handle* h, int i);
int64_t oard_weight_numel(const oard_handle* h, int i);

/* Copy one named fp32 tensor (row-major, exactly oard_weight_numel elements) into the handle.
 * `is_device` != 0: `data` is a device pointer on the handle's device; else a host pointer. */
int oard_set_weight(oard_handle* h, const char* name, const float* data, int64_t numel, int is_device, void* stream);
/* Verify that every required weight was set; derive packed forms.  Must be called after (re)setting weights. */
int oard_commit_weights(oard_handle* h, void* stream);

/* Build the static per-batch plan from the HOST edge list edge_index[2][E] (int64, edge_index[0] = source):
 * CSR rows, transposed-edge map, connected components (= reactions).  Requirements: edges grouped by source in
 * non-decreasing order, the graph symmetric, no duplicates, components of <= 256 nodes.  Allocates the workspace; a later
 * plan on the same handle (a sampler plans once per trajectory) keeps the buffers and grows only those that are too small, so
 * the workspace stays at its high-water mark until oard_destroy (oard_workspace_bytes reports the allocated capacity).
 * Components need not be complete graphs: when every component is complete (what get_edges_index builds per sample) the
 * message aggregation takes the group-staged kernel, otherwise the node-per-block kernel, which assumes nothing. */
int oard_plan(oard_handle* h, int64_t n_nodes, int64_t n_edges, const int64_t* edge_index_host);
size_t oard_workspace_bytes(const oard_handle* h);

/* LEFTNet.forward (leftnet.py:724-891).  h_in[N,C], pos[N,3], subgraph_mask[E] (int64, may be NULL = all ones),
 * h_out[N,C], dpos[N,3] (the reference returns pos + dpos).  Device pointers; asynchronous on `stream`.
 * subgraph_mask is in the order of the planned edge list.  On complete components it must be what
 * get_subgraph_mask (_graph_tools.py:39-59) produces from per-node fragment ids, i.e. an equivalence relation — the
 * group-staged message kernel relies on that (any other 0/1 mask: run with OARD_EQUI=node). */
int oard_forward(oard_handle* h, const float* h_in, const float* pos, const int64_t* subgraph_mask, float* h_out,
                 float* dpos, void* stream);

/* ---- Device-resident dynamics wrapper and reverse-diffusion step ----
 *   oa_reactdiff/dynamics/egnn_dynamics.py:63-168   EGNNDynamics.forward        -> oard_dyn_forward
 *   oa_reactdiff/dynamics/_base.py:82-132           encoders / decoders         -> oard_dyn_configure, oard_dyn_set_weight
 *   oa_reactdiff/diffusion/en_diffusion.py:562-632  sample_p_zs_given_zt        -> oard_reverse_step
 * The per-fragment encoder MLP, the time/condition channels, the NaN guard, the per-(fragment, sample) centre-of-mass
 * removal, the decoder MLP and (for the step) the posterior mean, noise projection and h0 overwrite run as three small
 * kernels around the LEFTNet forward; the whole call is replayed as ONE CUDA graph keyed by the caller's pointers
 * (callers keep persistent buffers).  Weight names are the reference state-dict names below `dynamics.`:
 * "encoders.{f}.mlp.{0,1}.linear.{weight,bias}", "decoders.{f}.mlp.{0,1}.linear.{weight,bias}". */
int oard_dyn_configure(oard_handle* h, int n_frag, int node_nf, int condition_nf, int condition_time);
int oard_dyn_num_weights(const oard_handle* h);
const char* oard_dyn_weight_name(const oard_handle* h, int i);
int64_t oard_dyn_weight_numel(const oard_handle* h, int i);
int oard_dyn_set_weight(oard_handle* h, const char* name, const float* data, int64_t numel, int is_device, void* stream);
/* After oard_plan: HOST arrays node_frag[N] (n_frag_switch, _graph_tools.py:62-81) and node_sample[N] (combined_mask,
 * :84-96).  The nodes of every (fragment, sample) pair must be contiguous (they are in the reference's node order). */
int oard_dyn_plan(oard_handle* h, const int64_t* node_frag_host, const int64_t* node_sample_host, int64_t n_samples);
/* eps[N, node_nf] = EGNNDynamics.forward(xh[N, node_nf] (fragments concatenated), t[B], conditions[B, condition_nf]).
 * Device pointers; t may be NULL when condition_time == 0, conditions when condition_nf == 0. */
int oard_dyn_forward(oard_handle* h, const float* xh, const float* t, const float* conditions,
                     const int64_t* subgraph_mask, float* eps, void* stream);
/* One reverse step z_t -> z_s IN PLACE on z[N, node_nf] with one (s, t) pair for the whole batch:
 *   eps = dynamics(z, t); mu = z / alpha_ts - eps * coef; z_s = mu + sigma * noise,
 * noise_pos[N, 3] / noise_feat[N, node_nf - 3] = the caller's raw standard-normal draws (the position noise is projected
 * on the zero-CoM subspace here, as are the new positions); noise_feat == NULL: zero feature noise (pos_only);
 * h0[N, node_nf - 3] != NULL: features overwritten by h0 (pos_only, en_diffusion.py:524-527).
 * t, alpha_ts, coef, sigma are the host-side schedule scalars of the step (diffusion/_schedule.py:132-203). */
int oard_reverse_step(oard_handle* h, float* z, const float* noise_pos, const float* noise_feat, const float* h0,
                      const float* conditions, const int64_t* subgraph_mask, float t, float alpha_ts, float coef,
                      float sigma, void* stream);
/* One RePaint step (en_diffusion.py:788-853, the body of inpaint()'s inner loop) IN PLACE on z: oard_reverse_step for the
 * fragments being generated, and for the clamped fragments f (bit f of known_frag_bits) a fresh draw from
 *   q(z_s | x_fixed) = alpha_s x_fixed + sigma_s eps'      (noised_representation, :260-279)
 * with eps' = noise_known_pos[N, 3] (projected on the zero-CoM subspace per (fragment, sample)) / noise_known_feat (NULL:
 * zero); then the h0 overwrite.  x_fixed[N, node_nf] holds the clamped fragments' data (other rows are ignored).  The whole
 * call is ONE CUDA-graph launch keyed by the pointers and known_frag_bits. */
int oard_inpaint_step(oard_handle* h, float* z, const float* noise_pos, const float* noise_feat, const float* h0,
                      const float* conditions, const int64_t* subgraph_mask, float t, float alpha_ts, float coef,
                      float sigma, const float* x_fixed, int known_frag_bits, const float* noise_known_pos,
                      const float* noise_known_feat, float alpha_s, float sigma_s, void* stream);
/* The RePaint jump-back z_s -> z_t (en_diffusion.py:1050-1074, sample_p_zt_given_zs) IN PLACE on z:
 *   z = alpha_ts z + sigma_ts eps, eps positions CoM-free per (fragment, sample), then the CoM of the new positions removed. */
int oard_jump_back(oard_handle* h, float* z, const float* noise_pos, const float* noise_feat, float alpha_ts,
                   float sigma_ts, void* stream);

/* ---- Training: differentiable forward + backward (BASELINE config 5) ----
 *   oa_reactdiff/model/leftnet.py:724-891 under torch autograd  -> oard_forward_train + oard_backward
 * oard_forward_train = LEFTNet.forward in exact fp32 with every activation the backward needs kept in the handle
 * (csrc/train_core.h; the graph artefacts and geometry come from the inference kernels: no parameter lies upstream of
 * the positions, so they carry no gradient).  oard_backward consumes dL/dh_out [N,C] and dL/ddpos [N,3] (device),
 * writes dL/dh_in [N,C] and ACCUMULATES dL/dparameter into per-weight gradient buffers (reference state-dict names),
 * read with oard_get_grad (device destination) and cleared with oard_zero_grads.  First, un-tuned version. */
int oard_forward_train(oard_handle* h, const float* h_in, const float* pos, const int64_t* subgraph_mask, float* h_out,
                       float* dpos, void* stream);
int oard_backward(oard_handle* h, const float* g_h_out, const float* g_dpos, float* g_h_in, void* stream);
int oard_zero_grads(oard_handle* h, void* stream);
int oard_get_grad(oard_handle* h, const char* name, float* dst_device, int64_t numel, void* stream);

/* Parity instrumentation.  With debug on, oard_forward keeps snapshots of intermediates; oard_debug_read copies a
 * named snapshot to host (synchronises).  Names: mask(u8[E]) group(i32[N]) act_idx(i32[n_act]) n_act(i32[1])
 * pos_frame(f32[N,3]) geo(f32[E,4]) rb f_act rbf_act s0 NE1 e0 nodeframe pos_prjt s_msg{l} vec_msg{l} e{l} s{l} vec{l}. */
int oard_set_debug(oard_handle* h, int on);
int64_t oard_debug_bytes(oard_handle* h, const char* name);
int oard_debug_read(oard_handle* h, const char* name, void* host_dst, size_t bytes);

/* Kernel launches issued by the last oard_forward call / by all calls so far (for bench.py's gpu_launches). */
int64_t oard_last_launch_count(const oard_handle* h);
int64_t oard_total_launch_count(const oard_handle* h);

/* Live per-kernel-class timing for the roofline: with every_n > 0, every every_n-th oard_forward brackets each launch
 * with CUDA events on the caller's stream and (synchronising at the end of that call) accumulates, per class tag,
 * device milliseconds, launches, algorithmic flops and algorithmic bytes.  every_n = 0 switches it off and clears. */
int oard_set_profile(oard_handle* h, int every_n);
int oard_profile_count(const oard_handle* h);
int oard_profile_get(const oard_handle* h, int i, const char** tag, double* ms, int64_t* launches, double* flops,
                     double* bytes);

/* Unit-test entry for the two GEMM implementations: C[M,N] = act(A[M,K] W[N,K]^T + bias), device pointers.
 * use_tc = 0: exact-fp32 SIMT kernel; 1: tcgen05 bf16x3 kernel (sm_100 only).  Synchronises `stream` when use_tc. */
int oard_test_gemm(int device, int M, int N, int K, const float* A, const float* W, const float* bias, float* C,
                   int use_tc, int act, int swap_lbo_sbo, void* stream);
/* Extended form: epilogue mode (0 plain, 1 two gathered adds from aux[M,2N], 2 multiply by aux[M,N], 3 residual
 * aux[M,N]), ablation bits for roofline studies (see gemm_tc.cuh TcDebugOpts), event-timed repetitions. */
int oard_test_gemm_ex(int device, int M, int N, int K, const float* A, const float* W, const float* bias, float* C,
                      int use_tc, int act, int swap_lbo_sbo, int mode, const float* aux, int ablate, int reps,
                      float* ms_out, void* stream);

/* Unit-test / timing entry of the pair16 GEMM (csrc/gemm_p16.cuh: A operand and, optionally, the output stored in
 * HBM as split-bf16 pairs in the tensor core's operand layout).  All tensors fp32 on the device; the entry converts.
 * mode as above (3 updates C in place); c2_out != NULL also returns the compact copy of rows m % 3 == 0;
 * ew: epilogue warps (0 = default, 8, 16). */
int oard_test_gemm_p16(int device, int M, int N, int K, const float* A, const float* W, const float* bias, float* C,
                       int mode, const float* aux, int out_pair, int act, float* c2_out, int ew, int reps,
                       float* ms_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OARD_H */
