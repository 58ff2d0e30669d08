/*
 * vbnn.h -- C ABI of libvbnn.so: the B200-native (sm_100a) VBLinear hot path of louissmit/VBNN.
 *
 * This is the drop-in boundary.  Every entry point replaces one call the reference makes
 * through Torch7's nn.Module protocol on class nn.VBLinear (reference VBLinear.lua:7) or through
 * the mlp.lua net object; the reference interface each one replaces is cited as file:line.
 * The reference-side binding (LuaJIT ffi.cdef over this header) is lua/VBLinear.lua; see
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain C, opaque handles, plain pointers and sizes; no C++/torch types cross the ABI;
 *   - every function returns int: 0 = VBNN_OK, negative = VBNN_E_*; vbnn_last_error() gives a
 *     thread-local message; nothing throws or exits across the ABI;
 *   - all tensors are fp32, row-major, dense, exactly as the reference's torch.FloatTensor /
 *     CudaTensor (main.lua:10): weights are [outputSize x inputSize] (VBLinear.lua:18-23),
 *     activations are [N x features]; class targets are 1-based floats (data.lua:16);
 *   - pointers named *_dev are device pointers on the context's GPU, borrowed for the call;
 *     pointers named *_host are host pointers (pinned memory makes the copies asynchronous);
 *   - work is enqueued on the context's stream; a function that returns a host scalar
 *     synchronises that stream, nothing else does;
 *   - there is no CPU fallback: without a CUDA device every compute entry returns VBNN_E_CUDA.
 */
#ifndef VBNN_H
#define VBNN_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VBNN_ABI_VERSION 2

typedef struct vbnn_ctx vbnn_ctx;     /* one per (GPU, host thread)                         */
typedef struct vbnn_layer vbnn_layer; /* one nn.VBLinear (VBLinear.lua:7) or plain nn.Linear */
typedef struct vbnn_mlp vbnn_mlp;     /* the mlp.lua net object (mlp.lua:5-143)              */

enum vbnn_status {
  VBNN_OK = 0,
  VBNN_E_INVALID = -1,     /* bad argument / shape mismatch (Torch would raise a Lua error) */
  VBNN_E_CUDA = -2,        /* CUDA runtime/driver error, or no device                       */
  VBNN_E_NOMEM = -3,
  VBNN_E_UNSUPPORTED = -4,
  VBNN_E_NCCL = -5,
  VBNN_E_STATE = -6        /* call order violated (e.g. backward before forward)            */
};

enum vbnn_reparam {
  VBNN_REPARAM_WEIGHT = 0, /* W = mu + sigma*eps per MC sample -- the reference (VBLinear.lua:59)  */
  VBNN_REPARAM_LOCAL = 1   /* local reparameterisation: Y = X mu^T + sqrt(X^2 s2^T)*zeta (new)    */
};

enum vbnn_precision {
  VBNN_PREC_FP32 = 0,      /* fp32 operands, fp32 accumulate, CUDA-core GEMM (exact-parity mode)   */
  VBNN_PREC_BF16 = 1       /* bf16 operands, fp32 accumulate, tcgen05/TMEM tensor-core GEMM        */
};

enum vbnn_kind { VBNN_KIND_VB = 0, VBNN_KIND_LINEAR = 1 };

/* names for vbnn_layer_get / vbnn_layer_set; all fp32, [O x I] unless noted */
enum vbnn_buf {
  VBNN_BUF_MEANS = 0,      /* self.means      VBLinear.lua:23                  */
  VBNN_BUF_LVARS = 1,      /* self.lvars      VBLinear.lua:18                  */
  VBNN_BUF_BIAS = 2,       /* self.bias [O]   VBLinear.lua:13                  */
  VBNN_BUF_WEIGHT = 3,     /* self.weight (last sampled W)   VBLinear.lua:63   */
  VBNN_BUF_GRAD_WEIGHT = 4,/* self.gradWeight VBLinear.lua:113                 */
  VBNN_BUF_GRAD_SUM = 5,   /* self.gradSum    VBLinear.lua:20,115              */
  VBNN_BUF_GRAD_BIAS = 6,  /* self.gradBias [O]                                */
  VBNN_BUF_ADAM_M_MU = 7,  /* meanState.m     VBLinear.lua:32,135              */
  VBNN_BUF_ADAM_V_MU = 8,
  VBNN_BUF_ADAM_M_VAR = 9, /* varState.m      VBLinear.lua:33,140              */
  VBNN_BUF_ADAM_V_VAR = 10,
  VBNN_BUF_EPS = 11,       /* self.e (last epsilon)  VBLinear.lua:37,55        */
  VBNN_BUF_STDV = 12,      /* self.stdv cached by compute_prior  VBLinear.lua:79 */
  VBNN_BUF_MU_SQE = 13,    /* self.mu_sqe cached by compute_prior VBLinear.lua:82 */
  VBNN_BUF__COUNT = 14
};

/* config.lua restated; vbnn_opts_default() fills the reference's shipped values. */
typedef struct vbnn_opts {
  float var_init;          /* config.lua:44 (0.001)                                        */
  int msr_init;            /* config.lua:45: var_init = 2/inputSize (VBLinear.lua:14-16)   */
  float mu_init;           /* config.lua:43: 0 -> means = 0, else means ~ N(0, var_init)   */
  float B;                 /* config.lua:30 (1e6): number of minibatches, scales the KL    */
  int S;                   /* config.lua:32: Monte-Carlo weight samples per minibatch      */
  float lr_bias;           /* opt.state.learningRate      config.lua:51-54 (1e-3)          */
  float lr_mu;             /* opt.meanState.learningRate  config.lua:60-64 (1e-4)          */
  float lr_var;            /* opt.varState.learningRate   config.lua:55-59 (0.05)          */
  float adam_beta1, adam_beta2, adam_eps; /* optim.adam defaults 0.9, 0.999, 1e-8         */
  int reparam;             /* enum vbnn_reparam                                            */
  int precision;           /* enum vbnn_precision                                          */
  int strict_reference;    /* 1: reproduce quirks Q1/Q6 (sample() and calc_lc() use the
                              sigma / mu^2 cached by the last compute_prior)               */
} vbnn_opts;

/* the 14 diagnostics VBLinear.lua:150-163 logs, in file order */
typedef struct vbnn_stats {
  float vlc_grad, vle_grad, mlc_grad, mle_grad;
  float min_variance, max_variance, mean_variance, var_hat;
  float mean_means, std_means, min_means, max_means;
  float mu_normratio, var_normratio;
} vbnn_stats;

/* ---------------------------------------------------------------- library ------------ */
int vbnn_abi_version(void);
const char* vbnn_last_error(void);
void vbnn_opts_default(vbnn_opts* opts);

/* ---------------------------------------------------------------- context ------------
 * replaces: require 'cunn' + the implicit cutorch default stream (VBLinear.lua:2, main.lua:1),
 * torch.manualSeed (config.lua:40).  `stream` is a cudaStream_t (NULL = a private stream). */
int vbnn_ctx_create(int device, void* stream, uint64_t seed, vbnn_ctx** out);
/* Which stream the context enqueues on.  The reference's other modules (nn.ReLU, nn.LogSoftMax, the
 * criterion: mlp.lua:19,27,30,32) run on cutorch's LEGACY DEFAULT stream (stream 0), with which a
 * non-blocking private stream does not synchronise -- a layer-level drop-in must therefore share it:
 *   VBNN_CTX_STREAM_GIVEN           `stream` as passed (NULL = private non-blocking stream: the net-level
 *                                   entry points, which own the whole minibatch; = vbnn_ctx_create)
 *   VBNN_CTX_STREAM_LEGACY_DEFAULT  stream 0 itself (`stream` must be NULL): every call is ordered with
 *                                   the caller's default-stream work exactly like a cunn module.  CUDA
 *                                   graphs cannot capture stream 0, so vbnn_mlp_step launches eagerly.
 *   VBNN_CTX_STREAM_PRIVATE_BLOCKING a private stream created without cudaStreamNonBlocking: it keeps
 *                                   graph replay and still synchronises implicitly with stream 0. */
enum vbnn_ctx_stream {
  VBNN_CTX_STREAM_GIVEN = 0,
  VBNN_CTX_STREAM_LEGACY_DEFAULT = 1,
  VBNN_CTX_STREAM_PRIVATE_BLOCKING = 2
};
int vbnn_ctx_create_ex(int device, void* stream, int stream_mode, uint64_t seed, vbnn_ctx** out);
int vbnn_ctx_destroy(vbnn_ctx* ctx);
int vbnn_ctx_synchronize(vbnn_ctx* ctx);
/* Per-launch device timing of the tensor-core GEMM: CUDA events on the context's stream around
 * every launch, summed per epilogue class (0 store, 1 fwd, 2 fwd-lrt, 3 dx, 4 dx-lrt, 5 dw,
 * 6 dw-lrt) together with the algorithmic flops; class 7 is the fused KL + Adam update, whose
 * "flops" slot holds its algorithmic HBM bytes (56 B per weight).  Enabling it disables graph replay.  This is
 * what bench.py's roofline object is computed from. */
int vbnn_ctx_profile(vbnn_ctx* ctx, int enable);
int vbnn_ctx_profile_read(vbnn_ctx* ctx, int cls, double* total_ms, long long* launches, double* flops);
/* same switch: time between phase marks of the minibatch (1 start..params ready, 2 sample, 3 forward,
 * 4 loss, 5 backward (per layer), 6 exchange + update, 7 finalise) */
int vbnn_ctx_phase_read(vbnn_ctx* ctx, int id, double* total_ms, long long* count);
int vbnn_ctx_set_step(vbnn_ctx* ctx, uint32_t step);   /* Philox "minibatch" counter */
int vbnn_ctx_get_step(vbnn_ctx* ctx, uint32_t* step);

/* ---------------------------------------------------------------- nn.VBLinear --------- */
/* nn.VBLinear(inputSize, outputSize, opt)  VBLinear.lua:9-47 (kind VB) or nn.Linear (mlp.lua:29) */
int vbnn_layer_create(vbnn_ctx* ctx, int inputSize, int outputSize, int kind,
                      const vbnn_opts* opts, vbnn_layer** out);
int vbnn_layer_destroy(vbnn_layer* layer);
int vbnn_layer_dims(const vbnn_layer* layer, int* inputSize, int* outputSize);

/* VBLinear:sample(opt)  VBLinear.lua:49-64.  eps_dev == NULL: epsilon is drawn on the device
 * (Philox4x32-10, counter = (element, layer, sample_idx, step)); otherwise the [O x I] tensor
 * eps_dev is injected (parity mode) and kept as self.e. */
int vbnn_layer_sample(vbnn_layer* layer, int sample_idx, const float* eps_dev);
/* VBLinear:clamp_to_map()  VBLinear.lua:105-107 */
int vbnn_layer_clamp_to_map(vbnn_layer* layer);

/* nn.Linear:updateOutput(input)  (inherited; invoked via mlp.lua:77).  Y_dev [N x O].
 * In VBNN_REPARAM_LOCAL mode zeta_dev ([N x O], nullable) injects the activation noise. */
int vbnn_layer_forward(vbnn_layer* layer, const float* X_dev, int N, float* Y_dev,
                       const float* zeta_dev);
/* nn.Linear:updateGradInput(input, gradOutput)  (inherited; via mlp.lua:79).  dX_dev [N x I]. */
int vbnn_layer_backward_data(vbnn_layer* layer, const float* X_dev, const float* G_dev, int N,
                             float* dX_dev);
/* VBLinear:accGradParameters(input, gradOutput, scale)  VBLinear.lua:112-118
 * (one GEMM, not the reference's two -- quirk Q2 -- with the same result). */
int vbnn_layer_acc_grad(vbnn_layer* layer, const float* X_dev, const float* G_dev, int N,
                        float scale);
/* VBLinear:resetAcc() VBLinear.lua:120-122 plus gradParameters:zero() (mlp.lua:63) */
int vbnn_layer_reset_acc(vbnn_layer* layer);

/* VBLinear:compute_prior()  VBLinear.lua:77-88 -> mu_hat, var_hat (host scalars, synchronises) */
int vbnn_layer_compute_prior(vbnn_layer* layer, float* mu_hat, float* var_hat);
/* VBLinear:compute_mugrads / compute_vargrads  VBLinear.lua:90-98.  Each out pointer is a
 * nullable [O x I] device buffer: mleg/vleg likelihood terms, mlcg/vlcg complexity terms.
 * Like the reference this divides gradWeight / gradSum in place. */
int vbnn_layer_grads(vbnn_layer* layer, float* mleg_dev, float* mlcg_dev, float* vleg_dev,
                     float* vlcg_dev);
/* VBLinear:update(opt)  VBLinear.lua:124-166.  stats (nullable) receives the 14 diagnostics
 * of VBLinear.lua:150-163 and synchronises; NULL keeps the call asynchronous. */
int vbnn_layer_update(vbnn_layer* layer, vbnn_stats* stats);
/* VBLinear:calc_lc(opt)  VBLinear.lua:99-103.  lc_dev: nullable [O x I]; sum_host: nullable. */
int vbnn_layer_calc_lc(vbnn_layer* layer, float* lc_dev, float* sum_host);

/* parameters() / field access (mlp.lua:37,48-49; main.lua:123; checkpointing utils.lua:73-80) */
int vbnn_layer_get(vbnn_layer* layer, int which, float* dst_host);
int vbnn_layer_set(vbnn_layer* layer, int which, const float* src_host);
int vbnn_layer_device_ptr(vbnn_layer* layer, int which, float** ptr_dev, size_t* count);
/* Adopt CALLER-OWNED device storage for weight / bias / gradWeight / gradBias (which = VBNN_BUF_WEIGHT,
 * _BIAS, _GRAD_WEIGHT, _GRAD_BIAS): Torch7's getParameters() (mlp.lua:37) re-flattens these four tensors
 * of every module into one new storage, so the module must follow them there instead of the other way
 * round.  The caller's buffer is adopted AS IS (Torch copies the old values into the flat storage itself,
 * and mlp.lua:48-54 then rewrites weight / bias there: the tensor is the parameter) and must stay valid until
 * the next bind or vbnn_layer_destroy; ptr_dev == NULL hands the buffer back to the library (contents kept).
 * Layers owned by a vbnn_mlp keep their gradients in the mlp's arena and refuse. */
int vbnn_layer_bind(vbnn_layer* layer, int which, float* ptr_dev);
/* optimiser step counters (meanState.t / varState.t / biasState.evalCounter) */
int vbnn_layer_get_t(vbnn_layer* layer, int* t);
int vbnn_layer_set_t(vbnn_layer* layer, int t);
/* signal-to-noise pruning mask of mainviz.lua:20-24: counts |mu|/sigma < thresh */
int vbnn_layer_snr_count(vbnn_layer* layer, float thresh, uint8_t* mask_dev, long long* count);

/* draw the exact epsilon [O x I] (or, for local reparameterisation, zeta [N x O] with
 * rows = N) that the fused kernels generate for (step, sample_idx) -- parity tests inject it
 * into the CPU oracle. */
int vbnn_layer_draw_noise(vbnn_layer* layer, uint32_t step, int sample_idx, int rows,
                          int row0, float* out_dev);

/* ---------------------------------------------------------------- mlp.lua -------------
 * MLP:buildModel(opt)  mlp.lua:7-60.  sizes = {input_size, hidden..., #classes}; hidden layers
 * are VBLinear+ReLU, the last is nn.Linear (mlp.lua:29) or, with vb_output, a VBLinear
 * (convnet.lua:30).  Ends in LogSoftMax + ClassNLLCriterion (mlp.lua:30-32). */
int vbnn_mlp_create(vbnn_ctx* ctx, const int* sizes, int n_sizes, int vb_output, int max_batch,
                    const vbnn_opts* opts, vbnn_mlp** out);
int vbnn_mlp_destroy(vbnn_mlp* mlp);
int vbnn_mlp_num_layers(const vbnn_mlp* mlp);
int vbnn_mlp_layer(vbnn_mlp* mlp, int k, vbnn_layer** out);   /* borrowed handle */
/* parameter init on the device: means per VBLinear.lua:22-29, Linear weights ~ N(0, 2/fan_in)
 * and zero biases per mlp.lua:47-55 */
int vbnn_mlp_init_params(vbnn_mlp* mlp, uint64_t seed, int he_means);

/* the step pieces, same names and order as mlp.lua / main.lua:28-40 */
int vbnn_mlp_reset_gradients(vbnn_mlp* mlp);                               /* mlp.lua:62-67  */
int vbnn_mlp_sample(vbnn_mlp* mlp, int sample_idx);                        /* mlp.lua:69-74  */
int vbnn_mlp_run(vbnn_mlp* mlp, const float* X_dev, const float* targets_dev, int N,
                 int sample_idx, float* err_host, float* acc_host);        /* mlp.lua:76-84  */
int vbnn_mlp_update(vbnn_mlp* mlp);                                        /* mlp.lua:117-142 */
int vbnn_mlp_calc_lc(vbnn_mlp* mlp, float* lc_host);                       /* mlp.lua:109-115 */

/* One whole minibatch of main.lua:28-40 (reset, S x (sample, run), [allreduce], update) as one
 * enqueued sequence (CUDA graph when the shape repeats).  result_dev: nullable device float[2]
 * = {mean error, mean accuracy %} as main.lua:38-39. */
int vbnn_mlp_step(vbnn_mlp* mlp, const float* X_dev, const float* targets_dev, int N,
                  float* result_dev);
/* same with HOST buffers: H2D of inputs + D2H of {error, accuracy} inside the call
 * (replaces inputs:cuda()/targets:cuda() main.lua:23-24 and the scalar reads mlp.lua:80-82). */
int vbnn_mlp_step_host(vbnn_mlp* mlp, const float* X_host, const float* targets_host, int N,
                       float* err_host, float* acc_host);
/* pipelined host-buffer variant: submit copies minibatch t+1 while minibatch t computes;
 * collect returns the {error, accuracy} of the oldest submitted minibatch. */
int vbnn_mlp_submit_host(vbnn_mlp* mlp, const float* X_host, const float* targets_host, int N);
int vbnn_mlp_collect(vbnn_mlp* mlp, float* err_host, float* acc_host);
/* The same pipeline fed with the dataset's NATIVE bytes: MNIST pixels are uint8 before data.lua:25,30
 * (u.normalize, utils.lua:29-35) turns them into (x - mean) / std floats on the host.  Here the bytes cross
 * PCIe (4x fewer than fp32) and the normalisation is fused into the operand-staging kernel on the device:
 * x = (float(byte) - mean) * inv_std, then exactly the path of vbnn_mlp_submit_host. */
int vbnn_mlp_submit_host_u8(vbnn_mlp* mlp, const uint8_t* X_host, const float* targets_host, int N,
                            float mean, float inv_std);
/* Make the context's stream wait for everything a minibatch enqueued on the library's auxiliary streams (the
 * peer-mode side stream: last shard update + operand push; the NCCL stream), so that an event recorded on the
 * context's stream afterwards covers the WHOLE minibatch.  Asynchronous; a no-op on one GPU. */
int vbnn_mlp_join_streams(vbnn_mlp* mlp);

/* net:test(input, target)  mlp.lua:86-107: n_samples == 0 -> clamp_to_map (quicktest),
 * else mean over n_samples sampled forward passes (no wasted backward). */
int vbnn_mlp_test(vbnn_mlp* mlp, const float* X_dev, const float* targets_dev, int N,
                  int n_samples, float* err_host, float* acc_host);
/* last forward's LogSoftMax output [N x C] (self.model.output, visualize.lua:97) */
int vbnn_mlp_get_outputs(vbnn_mlp* mlp, int sample_idx, float* logp_host);
/* kernels launched by this handle since creation (bench.py's gpu_launches) */
int vbnn_mlp_launch_count(vbnn_mlp* mlp, long long* count);
/* flat gradient arena {gW, gS, gb per layer}: what the data-parallel allreduce sums */
int vbnn_mlp_grad_arena(vbnn_mlp* mlp, float** ptr_dev, size_t* count);

/* ---------------------------------------------------------------- data parallel -------
 * New functionality (the reference is single-GPU): one rank per GPU,
 * minibatch rows sharded, sum-allreduce of the gradient arena over NCCL/NVLink. */
int vbnn_comm_unique_id(void* id128);                       /* 128-byte ncclUniqueId       */
int vbnn_comm_init(vbnn_ctx* ctx, const void* id128, int rank, int nranks);
int vbnn_comm_destroy(vbnn_ctx* ctx);
int vbnn_comm_allreduce(vbnn_ctx* ctx, float* buf_dev, size_t count);

/* Peer mode: the same exchange without any collective call, fused into the hot path over NVLink
 * peer memory (CUDA IPC; one process per GPU of one box).  Rows [q*rpo, (q+1)*rpo) of every layer
 * belong to rank q: the dW GEMM epilogue stores each gradient tile straight into its owner's
 * receive slot (reduce-scatter), the owner's fused update sums the slots for its rows only, and
 * the copy engines push the refreshed operands to every rank (all-gather) while backward
 * continues.  Set-up: after vbnn_comm_init, every rank exports a blob (blob == NULL queries its
 * size), the host all-gathers the blobs (rank order, blob_len bytes each) and every rank imports
 * them.  vbnn_mlp_step / vbnn_mlp_submit_host then use the peer path. */
int vbnn_mlp_peer_export(vbnn_mlp* mlp, void* blob, size_t capacity, size_t* blob_len);
int vbnn_mlp_peer_import(vbnn_mlp* mlp, const void* blobs_all_ranks, size_t blob_len);
int vbnn_mlp_peer_active(const vbnn_mlp* mlp);
/* collective: refresh the fp32 state of rows owned by other ranks (Adam moments, and whatever training does
 * not push) before get / checkpoint / clamp_to_map.  Protocol, on EVERY rank: drain the device
 * (cudaDeviceSynchronize: the last minibatch's shard updates run on the library's side stream) -> host barrier
 * -> vbnn_mlp_sync_replicas -> host barrier.  A rank that pulls before its peer has drained reads rows one
 * optimiser step behind. */
int vbnn_mlp_sync_replicas(vbnn_mlp* mlp);
/* shard of an O-row matrix owned by `rank`; returns rows per owner (a multiple of 32) */
int vbnn_peer_shard(int O, int nranks, int rank, int* row0, int* rows);

/* ---------------------------------------------------------------- self-test hooks -----
 * Raw GEMM entry used by tests/profiling to exercise the tcgen05 kernel in isolation:
 * D[M x N] (fp32) = A * B with bf16 operands. a_kmajor: A is [M x K] row-major (1) or
 * [K x M] row-major (0); b_kmajor: B is [N x K] row-major (1) or [K x N] row-major (0). */
int vbnn_gemm_bf16(vbnn_ctx* ctx, const uint16_t* A_dev, int lda, int a_kmajor,
                   const uint16_t* B_dev, int ldb, int b_kmajor, float* D_dev, int ldd,
                   int M, int N, int K, int batch, long long strideA, long long strideB,
                   long long strideD);
int vbnn_philox_normal(vbnn_ctx* ctx, uint64_t seed, uint32_t step, uint32_t stream,
                       uint32_t sample, int rows, int cols, int row0, float* out_dev);
/* Experiment switches (csrc/knobs.h: "tc_bn", "tc_cg", "lrt_split", "dw_split", "no_graph", ...; the same
 * names upper-cased with a VBNN_ prefix are read from the environment once per process).  The parity
 * tests use this to force the CTA-pair 256 x 256 tiles / a GEMM form at oracle-sized problems.
 * value == INT_MIN restores the default; old_value is nullable.  Unknown name: VBNN_E_INVALID. */
int vbnn_debug_knob(const char* name, int value, int* old_value);

#ifdef __cplusplus
}
#endif
#endif /* VBNN_H */
