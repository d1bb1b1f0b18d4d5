/*
 * molann_b200.h -- C ABI of the B200-native molann hot path (align -> features -> MLP, fwd and d/dx).
 *
 * The reference (zwpku/molann) is pure Python on PyTorch and has NO FFI of its own; the boundary it
 * fixes is its Python class API (molann/ann.py).  This header is the layer UNDER that API: each entry
 * point replaces the ATen op sequence of one reference method and is what a binding (ctypes / cgo /
 * JNI / the torch custom-op shim in molann_b200/csrc/torch_shim.cpp) calls.  See INTEGRATION.md.
 *
 *   molann_b200_align_forward / _backward        <- AlignmentLayer.forward      molann/ann.py:157-199 (+ autograd)
 *   molann_b200_preprocess_forward / _backward   <- PreprocessingANN.forward    molann/ann.py:553-565
 *                                                   (= FeatureLayer.forward :454-474 when n_align == 0,
 *                                                    FeatureMap.forward :288-356 for a one-feature plan)
 *   molann_b200_forward / _backward              <- MolANN.forward              molann/ann.py:620-624
 *                                                   with ann_layers = create_sequential_nn(...) :37-67
 *
 * Conventions: plain C types; every function returns an int status (0 = MOLANN_OK), never throws;
 * all data pointers are DEVICE pointers owned by the caller unless the name ends in `_host`; no hidden
 * allocation (scratch comes from the caller-provided workspace); `stream` is a cudaStream_t passed as
 * void*; calls are asynchronous on that stream and re-entrant from any host thread.
 * Layouts: x [L, n_inp, 3] fp32 contiguous; features [L, d_feat]; y [L, dims[n_layers]];
 * W_k [dims[k+1], dims[k]] row-major (torch.nn.Linear.weight), b_k [dims[k+1]].
 */
#ifndef MOLANN_B200_H_
#define MOLANN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOLANN_B200_VERSION 100
#define MOLANN_MAX_LAYERS 8

/* feature type ids: molann/feature.py:87-97 */
#define MOLANN_FEAT_ANGLE 0
#define MOLANN_FEAT_BOND 1
#define MOLANN_FEAT_DIHEDRAL 2
#define MOLANN_FEAT_POSITION 3

/* activation ids (create_sequential_nn's `activation`, molann/ann.py:37,64) */
#define MOLANN_ACT_TANH 0
#define MOLANN_ACT_RELU 1
#define MOLANN_ACT_SIGMOID 2
#define MOLANN_ACT_IDENTITY 3

/* status codes */
#define MOLANN_OK 0
#define MOLANN_ERR_NULL 1        /* required pointer is NULL */
#define MOLANN_ERR_PLAN 2        /* inconsistent plan (sizes, layer count, activation id ...) */
#define MOLANN_ERR_WORKSPACE 3   /* workspace too small / NULL */
#define MOLANN_ERR_ALIGNMENT 4   /* pointer not 4-byte aligned */
#define MOLANN_ERR_CUDA 5        /* a CUDA runtime call failed (see molann_b200_last_cuda_error) */
#define MOLANN_ERR_UNSUPPORTED 6 /* valid request this build cannot serve (e.g. no MLP in plan) */

/* One feature-program entry = 6 int32: {type, a0, a1, a2, a3, out_col}.  Atom indices are LOCAL
 * (positions inside the input atom group, molann/ann.py:144,261).  A reference 'position' feature over
 * m atoms is expanded into m entries of type 3 with a0 = atom and out_col = first of its 3 columns
 * (atom-major, xyz-minor: molann/ann.py:354).  Unused atom slots are 0. */
#define MOLANN_ENTRY_INTS 6

typedef struct MolannPlan {
  int32_t n_inp;                 /* atoms per frame (input_atom_num, molann/ann.py:133)            */
  int32_t n_align;               /* alignment selection size; 0 = no AlignmentLayer (nn.Identity)  */
  const int32_t* align_idx;      /* [n_align] local indices (_local_align_atom_indices, :144)      */
  const float* ref_x;            /* [n_align*3] centred reference (buffer ref_x, :137-141)         */
  int32_t n_entries;             /* feature-program entries                                        */
  const int32_t* entries;        /* [n_entries*MOLANN_ENTRY_INTS]                                  */
  int32_t d_feat;                /* FeatureLayer.output_dimension(), :446-452                      */
  int32_t use_angle_value;       /* FeatureLayer use_angle_value, :253                             */
  int32_t n_layers;              /* number of Linear layers; 0 = preprocessing only               */
  int32_t act_id;                /* MOLANN_ACT_*: applied after every layer but the last (:62-65)  */
  int32_t dims[MOLANN_MAX_LAYERS + 1]; /* dims[0] == d_feat                                        */
  const float* W[MOLANN_MAX_LAYERS];
  const float* b[MOLANN_MAX_LAYERS];
} MolannPlan;

int molann_b200_version(void);
const char* molann_b200_strerror(int status);
/* last cudaError_t seen by this library on the calling thread's most recent failing call (0 if none) */
int molann_b200_last_cuda_error(void);
const char* molann_b200_cuda_error_string(int cuda_error);
/* number of kernels this library has launched since load (all threads) -- bench.py's gpu_launches */
int64_t molann_b200_launch_count(void);

/* Host-side consistency check of the scalar fields (device arrays are not dereferenced). */
int molann_b200_plan_validate(const MolannPlan* plan);

/* Scratch bytes needed by molann_b200_forward (want_backward == 0) or _backward (!= 0) for L frames.
 * May be 0 (the fused small-system kernels need no scratch). */
size_t molann_b200_workspace_bytes(const MolannPlan* plan, int64_t L, int want_backward);

/* y[L, dims[n_layers]] = MLP(features(align(x))) */
int molann_b200_forward(const MolannPlan* plan, const float* x, int64_t L, float* y,
                        void* workspace, size_t workspace_bytes, void* stream);

/* gx[L, n_inp, 3] = d<gy, y>/dx (overwritten, dense).  gW / gb: NULL, or arrays of n_layers device
 * pointers into which d<gy,y>/dW_k, /db_k are ACCUMULATED (+=; caller zeroes them).  gx may be NULL when
 * parameter gradients are requested (training: the coordinates do not require a gradient) -- the first layer's
 * input contraction and the whole preprocessing backward are then skipped. */
int molann_b200_backward(const MolannPlan* plan, const float* x, const float* gy, int64_t L, float* gx,
                         float* const* gW, float* const* gb,
                         void* workspace, size_t workspace_bytes, void* stream);

/* Biasing-force entry point: y = model(x) AND gx = d<gy,y>/dx in one pass over x (one fused kernel when the
 * plan is eligible, else forward + backward).  Not in the reference tree; it is what an MD-engine plugin that
 * loads the exported model does with two autograd calls (README.rst:51, molann/ann.py:109-111). */
int molann_b200_value_and_grad(const MolannPlan* plan, const float* x, const float* gy, int64_t L, float* y,
                               float* gx, void* workspace, size_t workspace_bytes, void* stream);

/* Jacobian / all-forces entry point: y[L, k] = model(x) and jac[k][L][n_inp][3] with
 * jac[o] = d y[:, o] / dx -- every output's gradient plane from ONE pass over x (plane-major so that each plane is a
 * dense [L, n_inp, 3] array like gx above).  What an MD plugin driving several collective variables needs per step;
 * the reference does it with one autograd call per output (README.rst:51, molann/ann.py:109-111).  One kernel launch
 * for the tensor-core small-system class, otherwise one value-and-gradient pass per output. */
size_t molann_b200_jacobian_workspace_bytes(const MolannPlan* plan, int64_t L);
int molann_b200_value_and_jacobian(const MolannPlan* plan, const float* x, int64_t L, float* y, float* jac,
                                   void* workspace, size_t workspace_bytes, void* stream);

/* Trajectory wire format (the reference only ever builds a frame as torch.tensor(ag.positions),
 * molann/ann.py:106): coordinates travel over PCIe as int16 steps of `resolution` around a batch origin, half the
 * bytes of fp32.  x[i] = origin_host[i % 3] + q[i] * resolution, one fp32 FMA (a host decoder doing the same FMA gets
 * the same bits, which is how the parity tests compare).  n_values = L * n_inp * 3.  `origin_host` is a HOST array. */
int molann_b200_decode_frames_i16(const int16_t* q, int64_t n_values, const float* origin_host, float resolution,
                                  float* x, void* stream);

/* feat[L, d_feat] = features(align(x)); the MLP fields of the plan are ignored */
int molann_b200_preprocess_forward(const MolannPlan* plan, const float* x, int64_t L, float* feat, void* stream);
int molann_b200_preprocess_backward(const MolannPlan* plan, const float* x, const float* gfeat, int64_t L,
                                    float* gx, void* stream);

/* out[L, n_inp, 3] = (x - c(x)) R(x); only n_inp, n_align, align_idx, ref_x of the plan are used */
int molann_b200_align_forward(const MolannPlan* plan, const float* x, int64_t L, float* out, void* stream);
int molann_b200_align_backward(const MolannPlan* plan, const float* x, const float* gout, int64_t L,
                               float* gx, void* stream);

/* ---- prepared plans (big systems with a wide first layer: BASELINE configs[2], [4]) -------------------------------
 * Everything that depends only on the plan -- the feature program regrouped in the kernel's operand order and the MLP
 * weights permuted, split into TF32 hi / lo and laid out per K-chunk for the tensor cores -- is built ONCE into a
 * caller-owned DEVICE buffer; afterwards MolANN.forward (molann/ann.py:620-624) is ONE persistent kernel per call with
 * no packing launches (csrc/fused_wide.cuh).  `MolannPrepared` is a small opaque HOST object describing that buffer.
 *
 *   molann_b200_prepared_bytes   device bytes the buffer needs (upper bound; 0: the plan is not eligible)
 *   molann_b200_prepare          builds it.  Copies the feature program to the host once (synchronises `stream`; do
 *                                not call while capturing a CUDA graph).  MOLANN_ERR_UNSUPPORTED if not eligible.
 *   molann_b200_prepared_refresh re-packs the weights after the caller changed them (asynchronous, two small kernels)
 *   molann_b200_prepared_workspace_bytes / molann_b200_forward_prepared   scratch size and the forward itself (the
 *                                fused wide kernel; the layered kernels on the packed operands when a frame does not
 *                                fit its shared-memory ring)
 *   molann_b200_prepared_destroy frees the host object (the device buffer stays the caller's)
 * The plan passed to the later calls must be the one prepared (same sizes; pointers may have moved only for W / b,
 * followed by a refresh). */
typedef struct MolannPrepared MolannPrepared;
/* 1 if molann_b200_forward_prepared is the kernel of choice for this plan (supported shape and no small-system
 * fused kernel applies; MOLANN_B200_WIDE = 0 / 1 forces never / whenever the shape is supported), else 0 */
int molann_b200_wide_eligible(const MolannPlan* plan);
size_t molann_b200_prepared_bytes(const MolannPlan* plan);
int molann_b200_prepare(const MolannPlan* plan, void* device_buffer, size_t bytes, void* stream, MolannPrepared** out);
int molann_b200_prepared_refresh(MolannPrepared* prepared, const MolannPlan* plan, void* stream);
size_t molann_b200_prepared_workspace_bytes(const MolannPrepared* prepared, int64_t L);
int molann_b200_forward_prepared(const MolannPrepared* prepared, const MolannPlan* plan, const float* x, int64_t L,
                                 float* y, void* workspace, size_t workspace_bytes, void* stream);
/* y and gx = d<gy, y>/dx on a prepared plan (replaces `y = model(x); torch.autograd.grad(y, x, gy)` of
 * molann/ann.py:553-565, 620-624): the forward is the fused wide kernel, which also leaves the hidden activations in the
 * workspace (tanh plans; other activations recompute the forward with the layered kernels); the backward runs the
 * layered tensor-core kernels on the operands packed once (no pack launches) and the block preprocess backward.  Frames
 * are processed in chunks of whole 148 x 128-frame waves that fit ~2 GB of workspace
 * (molann_b200_prepared_workspace_bytes covers it). */
int molann_b200_value_and_grad_prepared(const MolannPrepared* prepared, const MolannPlan* plan, const float* x,
                                        const float* gy, int64_t L, float* y, float* gx, void* workspace,
                                        size_t workspace_bytes, void* stream);
void molann_b200_prepared_destroy(MolannPrepared* prepared);

/* ---- autoencoder training step (BASELINE configs[3]) ------------------------------------------
 * The reference has no training loop; what its modules allow is the composition
 *   loss = mean((decoder(encoder(x)) - preprocessing(x))^2)
 * with encoder = MolANN (molann/ann.py:567-624), decoder = create_sequential_nn (ann.py:37-67) and the encoder's own
 * PreprocessingANN output (ann.py:553-565) as the target, differentiated by autograd w.r.t. every Linear parameter.
 * molann_b200_train_loss_and_grads replaces that whole graph (forward, loss, backward) by ONE fused kernel per call
 * plus a fixed-order reduction of the per-SM partial sums (csrc/fused_train.cuh):
 *   flat[0 .. P)  = d loss / d parameters in torch parameter order: encoder W_1, b_1, W_2, b_2, ... then decoder
 *                   W_1, b_1, ...  (each W row-major [out, in] like torch.nn.Linear.weight)
 *   flat[P]       = loss_scale * sum over frames and feature columns of (reconstruction - target)^2
 * `loss_scale` = 1 / (frames of the GLOBAL batch * d_feat) makes the sum over ranks of `flat` the gradient of the mean
 * over the global batch: one sum-allreduce of P + 1 floats is the only collective of a data-parallel step.
 * `dec->dims[0]` must equal the encoder's output width and the decoder's last width d_feat.  Deterministic: the same
 * inputs give the same bits.  molann_b200_sgd_apply is the plain SGD update p -= lr * g over the same flat layout. */
typedef struct MolannDecoder {
  int32_t n_layers;
  int32_t act_id;                          /* MOLANN_ACT_*, after every layer but the last */
  int32_t dims[MOLANN_MAX_LAYERS + 1];
  const float* W[MOLANN_MAX_LAYERS];
  const float* b[MOLANN_MAX_LAYERS];
} MolannDecoder;
/* 1 if the fused training kernel serves this pair (every activation row of one 128-frame tile and all weights fit the
 * SM's shared memory), else 0 */
int molann_b200_train_eligible(const MolannPlan* encoder, const MolannDecoder* decoder);
/* P: number of trainable floats (0 for an invalid pair) */
size_t molann_b200_train_param_count(const MolannPlan* encoder, const MolannDecoder* decoder);
size_t molann_b200_train_workspace_bytes(const MolannPlan* encoder, const MolannDecoder* decoder);
int molann_b200_train_loss_and_grads(const MolannPlan* encoder, const MolannDecoder* decoder, const float* x, int64_t L,
                                     float loss_scale, float* flat, void* workspace, size_t workspace_bytes,
                                     void* stream);
/* params[i][j] -= lr * flat[offset_i + j], offset_i = numel[0] + ... + numel[i-1]; n_params <= 4 * MOLANN_MAX_LAYERS.
 * `params` and `numel` are HOST arrays of device pointers / element counts. */
int molann_b200_sgd_apply(float* const* params, const int64_t* numel, int32_t n_params, const float* flat, float lr,
                          void* stream);

/* Data-parallel step without NCCL: one-shot sum-allreduce of `flat_local` (n floats: the result of
 * molann_b200_train_loss_and_grads on this rank's shard) over NVLink peer memory, fused with the SGD update of this rank's
 * parameters, in ONE kernel per rank.  `peer_buffers[r]` (r < world, HOST array) is the device address, valid on THIS
 * device, of rank r's symmetric buffer (CUDA IPC / torch symmetric memory) of molann_b200_allreduce_buffer_bytes(n, world)
 * bytes, zero-filled before the first step; `state` = 3 uint32 in this rank's device memory, initialised {1, 0, 0}.
 * Every rank must call it once per step with the same n; the sums are formed in rank order, so all ranks get the same
 * bits in `flat_global` and in their parameters.  lr == 0 leaves the parameters alone (allreduce only).
 * Replayable in a CUDA graph (the step counter lives in `state`).  world <= 8. */
size_t molann_b200_allreduce_buffer_bytes(int64_t n, int32_t world);
int molann_b200_allreduce_sgd(const float* flat_local, float* flat_global, int64_t n, void* const* peer_buffers,
                              int32_t rank, int32_t world, uint32_t* state, float* const* params,
                              const int64_t* numel, int32_t n_params, float lr, void* stream);

/* Tuning / introspection: which kernel family the dispatcher picks for this plan.
 * 0 = general (warp-per-frame geometry + layered GEMMs), 1 = fused small-system kernel. */
int molann_b200_path_for(const MolannPlan* plan, int want_backward);
/* Finer: 0 = general, 1 = fused FFMA kernel, 2 = fused tcgen05 (3xTF32 tensor-core MLP) kernel. */
int molann_b200_kernel_family(const MolannPlan* plan, int want_backward);

#ifdef __cplusplus
}
#endif
#endif /* MOLANN_B200_H_ */
