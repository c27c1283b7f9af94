/*
 * mode_engine.h — C ABI of the B200-native MoDE denoising engine (libmode_engine.so).
 *
 * Scope: the EDM/DDIM denoising loop over the MoDE transformer of intuitive-robots/MoDE_Diffusion_Policy,
 *   MoDEAgent.denoise_actions -> sample_ddim -> GCDenoiser.forward -> MoDeDiT.forward -> n_layers x NoiseBlockMoE.
 * The reference has no FFI for this path (it is first-party Python over ATen); its plug-in seam is the pair of Hydra
 * `_target_` strings in conf/model/mode_agent.yaml:41 and :47. Each entry point below names the reference function it
 * replaces (paths relative to the reference checkout). The Python mirror in mode_diffusion_policy_b200/ binds these with
 * ctypes; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions
 *   - return 0 on success, negative on error; mode_last_error() returns the thread-local message of the last failure
 *   - the caller owns every I/O buffer; the engine owns packed bf16 weights, workspace and CUDA graphs
 *   - `*_dev` pointers are fp32 device pointers on the engine's device; `*_host` are host pointers
 *   - work is enqueued on the caller's stream (cudaStream_t passed as void*); no hidden synchronisation except in
 *     the *_host entry points, mode_finalize_weights and the small getters, which say so
 *   - one engine per (device, stream) at a time; not thread-safe; independent engines may run concurrently
 *   - there is NO CPU fallback: every entry point fails with MODE_ERR_CUDA if no sm_100 device is present
 */
#ifndef MODE_ENGINE_H_
#define MODE_ENGINE_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mode_engine mode_engine_t;

enum {
  MODE_OK = 0,
  MODE_ERR_INVALID = -1,     /* bad argument / unsupported configuration */
  MODE_ERR_CUDA = -2,        /* CUDA runtime or driver failure (message holds the CUDA error string) */
  MODE_ERR_STATE = -3,       /* call order violated (e.g. forward before mode_finalize_weights) */
  MODE_ERR_UNKNOWN_NAME = -4 /* mode_set_weight: not a MoDeDiT state_dict key */
};

/* Constructor arguments of MoDeDiT (mode/models/networks/modedit.py:643-674) that shape the hot path, plus the
 * engine's capacity. Unsupported reference options (use_proprio, use_custom_attn_mask, use_shared_expert,
 * goal_conditioned=False, use_noise_token_as_input=False, linear_output=False) are rejected at create time. */
typedef struct mode_config {
  int32_t obs_dim;        /* width of each state_images token                        (conf/model/mode_agent.yaml:48) */
  int32_t goal_dim;       /* width of the language/goal embedding                    (:49) */
  int32_t action_dim;     /* <= 8                                                    (:51) */
  int32_t embed_dim;      /* d, multiple of 256, <= 2048                             (:56) */
  int32_t n_layers;       /* (:59) */
  int32_t n_heads;        /* head dim = d / n_heads must be 32, 64 or 128            (:60) */
  int32_t n_state_tokens; /* tokens in states['state_images'] (2 cameras in CALVIN, mode_agent.py:562-565) */
  int32_t action_seq_len; /* (:67) */
  int32_t num_experts;    /* <= 32                                                   (:69) */
  int32_t top_k;          /* <= 8, <= num_experts                                    (:70) */
  int32_t router_normalize; /* RouterCond.normalize (modedit.py:418-419)             */
  int32_t max_batch;      /* capacity: largest B any call will use */
  float sigma_data;       /* GCDenoiser.sigma_data (score_wrappers.py:26; conf: 0.5) */
  float rms_eps;          /* RMSNorm eps used by every norm in the blocks: 1e-6 (modedit.py:447, :470, :720) */
} mode_config_t;

/* Replaces MoDeDiT.__init__ (modedit.py:643-739): allocates packed-weight storage and workspace on the current
 * CUDA device. */
int mode_create(const mode_config_t* cfg, mode_engine_t** out);
void mode_destroy(mode_engine_t* e);
const char* mode_last_error(void);

/* Replaces nn.Module.load_state_dict for MoDeDiT: `name` is a reference state_dict key ("blocks.3.attn.key.weight",
 * "sigma_emb.bias", ...; the full list is in DESIGN.md), `data` is contiguous fp32 with the reference shape, on the
 * host (is_device = 0) or on the engine's device (1). The engine converts to its own layout (bf16, packed QKV,
 * interleaved SwiGLU rows); the caller keeps ownership. "gripper_embed.weight" is accepted and ignored (unused unless
 * use_proprio, modedit.py:684). Host sources are copied synchronously; device sources are packed by a kernel on the
 * default stream without synchronisation (mode_finalize_weights synchronises that stream). */
int mode_set_weight(mode_engine_t* e, const char* name, const void* data, int is_device, const int64_t* shape,
                    int ndim);
/* Verifies every tensor was provided, precomputes the sigma-embedding and router affine forms. Synchronous. */
int mode_finalize_weights(mode_engine_t* e);
/* The same two calls with the packing kernels (and the "weights ready" event) on `stream` — the stream on which the
 * caller's parameter updates were enqueued — instead of the legacy default stream. The plain entries above are these
 * with stream = NULL. A repack is additionally ordered after the engine's own earlier work only if that work ran on the
 * same stream (the Python layer always passes torch's current stream to both). */
int mode_set_weight_on_stream(mode_engine_t* e, const char* ref_name, const void* data, int is_device,
                              const int64_t* shape, int ndim, void* stream);
int mode_finalize_weights_on_stream(mode_engine_t* e, void* stream);

/* MoDeDiT.forward(states, actions, goals, sigma) (modedit.py:741-809), eval mode: raw network output F.
 * state_dev (B, n_state_tokens, obs_dim); goal_dev (B, goal_dim); actions_dev, out_dev (B, action_seq_len, action_dim);
 * sigma_dev (B,) or a single value when sigma_stride == 0. */
int mode_forward(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                 const float* sigma_dev, int sigma_stride, float* out_dev, int B, void* stream);

/* GCDenoiser.forward (score_wrappers.py:65-80): D(x; sigma) = c_out * F(c_in * x; sigma) + c_skip * x. */
int mode_denoise(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                 const float* sigma_dev, int sigma_stride, float* out_dev, int B, void* stream);

/* GCDenoiser.loss forward (score_wrappers.py:45-63), eval-mode routing, no dropout: noised = action + noise * sigma;
 * writes the scalar mean squared error to loss_dev[0] and the model output F to out_dev (may be NULL). */
int mode_loss(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* action_dev,
              const float* noise_dev, const float* sigma_dev, float* loss_dev, float* out_dev, int B, void* stream);

/* One training step of GCDenoiser.loss (score_wrappers.py:45-63 as called by MoDEAgent.diffusion_loss,
 * mode_agent.py:659-672) with the hand-written backward, deterministic mode (dropout 0, top-k routing):
 * forward with saved activations, loss -> loss_dev[0], model output F -> out_dev (may be NULL), and the gradient of the
 * mean loss w.r.t. every MoDeDiT parameter into the engine-owned flat fp32 gradient buffer. sigma_dev is per-sample
 * (B,). Un-routed experts receive exact zeros. Enqueued on `stream`; the first call allocates the training state. */
int mode_train_step(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* action_dev,
                    const float* noise_dev, const float* sigma_dev, float* loss_dev, float* out_dev, int B,
                    void* stream);
/* Stochastic regularisation of the reference's train mode for the following mode_train_step calls (all off by
 * default = the deterministic mode above):
 *   attn_pdrop   dropout on the attention probabilities (SDPA dropout_p, modedit.py:149; conf: 0.3)
 *   mlp_pdrop    nn.Dropout between SwishGLU and the down projection of every expert (modedit.py:254; conf: 0.1)
 *   goal_drop    MoDeDiT.mask_cond: each goal feature zeroed with this probability, no rescale (modedit.py:882-893; 0.1)
 *   embed_pdrop  nn.Dropout on the goal / image / action token embeddings (+ positions), modedit.py:779-784 (conf: 0)
 *   multinomial  use_argmax=False: every token draws its top_k experts with torch.multinomial(probs, k, False) semantics
 *                (modedit.py:389-390) instead of arg-max; routing becomes per token.
 * The random bits are a pure function of (seed, step, stream, layer, logical coordinates) (csrc/rng.cuh, restated in
 * oracle/mode_rng.py): `step` is the position of the next mode_train_step and advances by one per stochastic step, so
 * a run is reproducible from (seed, step) and the same masks can be handed to the reference for parity tests. */
int mode_train_set_stochastic(mode_engine_t* e, float attn_pdrop, float mlp_pdrop, float goal_drop, float embed_pdrop,
                              int multinomial, unsigned long long seed, unsigned int step);
/* Token-level routing of the last stochastic (multinomial) training step for `layer`: expert indices in draw order and
 * their renormalised probabilities, [B*T, top_k] each (host buffers, either may be NULL). Synchronises the device. */
int mode_train_get_token_routing(mode_engine_t* e, int layer, int B, int32_t* idx_host, float* w_host);
/* The flat gradient buffer (device pointer, element count): ONE all-reduce over it synchronises data-parallel ranks. */
int mode_grad_buffer(mode_engine_t* e, float** grads_dev, int64_t* numel);
/* Element offset and size of the gradient of the reference parameter `name` inside the flat buffer (reference tensor
 * layout, contiguous). */
int mode_grad_offset(mode_engine_t* e, const char* name, int64_t* offset, int64_t* numel);
/* Gradients of the last mode_train_step's loss w.r.t. its inputs: dstate_dev (B, n_state_tokens, obs_dim) and dgoal_dev
 * (B, goal_dim), fp32, either may be NULL. In the reference these flow on into the FiLM-ResNet encoders that produce
 * perceptual_emb (mode_agent.py:405-411, compute_input_embeddings :513-582); bf16 tensor-core operands like the rest. */
int mode_train_input_grads(mode_engine_t* e, float* dstate_dev, float* dgoal_dev, int B, void* stream);
/* Fused optimizer: AdamW with torch.optim.AdamW's arithmetic over the flat gradient buffer, fused
 * with the re-pack of the updated weights into the engine's GEMM layouts. The reference configures AdamW in
 * MoDEAgent.configure_optimizers / get_optim_groups (mode_agent.py:267-301, :362-384): lr, betas, eps, and weight decay
 * on every parameter whose name contains none of 'bias', 'LayerNorm', 'embedding'.
 * mode_optimizer_bind: `param_dev` is the caller-owned fp32 master of reference parameter `name` (reference layout,
 *   contiguous); it is updated in place by every step. weight_decay != 0 puts it in the decayed group. Parameters that
 *   are never bound (frozen routers, mode_agent.py:762-765) are left untouched. Re-binding with a new pointer is allowed.
 * mode_adamw_step: one launch over all bound tensors on `stream`, using the gradients of the last mode_train_step
 *   (all-reduced by the caller in data-parallel runs); `step` counts from 1 (bias correction); grad_scale_dev is an
 *   optional device scalar multiplied into every gradient (the loss tensor's incoming gradient). Moment buffers are
 *   engine-owned (mode_optimizer_state exposes them for checkpointing; gradient-buffer layout). */
int mode_optimizer_bind(mode_engine_t* e, const char* name, float* param_dev, int weight_decay);
int mode_optimizer_unbind_all(mode_engine_t* e);
int mode_adamw_step(mode_engine_t* e, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                    const float* grad_scale_dev, void* stream);
/* The same update issued in n_layers + 1 launches so that a data-parallel caller can pipeline it with the gradient
 * exchange: group l (0..n_layers-1) = block l's large tensors (>= 2^20 elements: q/k/v/c_proj, expert up/down, router
 * W1 — exactly the per-layer buckets of parallel.GradAllReduce), group n_layers = every remaining tensor. Each group
 * may go to its own stream as soon as its gradients are final; group n_layers must come last (it refreshes the derived
 * weights). Launching all groups once equals one mode_adamw_step, bit for bit. */
int mode_adamw_step_group(mode_engine_t* e, float lr, float beta1, float beta2, float eps, float weight_decay, int step,
                          const float* grad_scale_dev, int group, void* stream);
int mode_optimizer_state(mode_engine_t* e, float** exp_avg_dev, float** exp_avg_sq_dev, int64_t* numel);
/* Sharded optimizer state for data-parallel training (replaces DDP's all-reduce + replicated torch.optim.AdamW,
 * reference mode/training_calvin.py:92-103 + mode_agent.py:267-301, with reduce-scatter -> 1/world of the update per
 * rank -> all-gather of the bf16 weights). After mode_optimizer_set_sharding(rank, world > 1):
 *   - group l (0..n_layers-1) holds block l's large tensors that can be split evenly (bf16-packed, 1024-element blocks
 *     divisible by world; mode_optimizer_shard_tensors lists their (offset, numel) spans of the flat buffers — the
 *     caller reduce-scatters exactly these, rank r keeping elements [r, r+1) * numel / world of each); every other
 *     tensor moves to group n_layers and stays replicated (all-reduced by the caller);
 *   - mode_adamw_step_group(l) updates only this rank's part of each tensor (fp32 master, moments, EMA) and writes the new
 *     weights as bf16 into the staging buffer (mode_optimizer_staging: gradient layout) instead of the packed copies;
 *   - after the caller all-gathered the staging spans, mode_optimizer_pack_group(l) writes every packed copy of the group.
 * The fp32 masters, moments and EMA of a sharded tensor are only current on the owning rank until the caller gathers
 * them (optim.EngineAdamW.synchronize_parameters / state_dict). world <= 1 switches sharding off. */
int mode_optimizer_set_sharding(mode_engine_t* e, int rank, int world);
int mode_optimizer_shard_tensors(mode_engine_t* e, int group, int64_t* offsets, int64_t* numels, int capacity, int* n);
int mode_optimizer_staging(mode_engine_t* e, void** staging_dev, int64_t* numel);
int mode_optimizer_pack_group(mode_engine_t* e, int group, void* stream);
/* Block `layer`'s packed weights become current when the work enqueued so far on `stream` has finished (the engine records
 * an event it owns). The next call that reads them waits on its own stream: mode_train_step block by block — so the
 * forward of block 0 starts while the sharded optimizer is still gathering blocks 1..n-1 — every other entry for all
 * pending blocks before its first launch. */
int mode_weights_record_ready(mode_engine_t* e, int layer, void* stream);
/* Exponential moving average of the bound parameters inside the optimizer launch (reference mode/callbacks/ema.py:
 * ema -= (1 - decay) * (ema - w) after every step, :119-126; +8 bytes per parameter instead of a separate pass over
 * all weights). decay in [0, 1] enables it for the following steps (may change every step: ema.py:84-92 warm-up
 * schedule), negative disables. The first enabled step seeds the average with the weights before the update (the
 * callback clones them at on_train_start, ema.py:96). mode_optimizer_ema_state exposes the engine-owned buffer
 * (gradient-buffer layout: mode_grad_offset gives each parameter's span). */
int mode_optimizer_set_ema(mode_engine_t* e, double decay);
int mode_optimizer_ema_state(mode_engine_t* e, float** ema_dev, int64_t* numel);
/* Checkpoint resume: the caller has copied a saved average into the buffer above; the next step must update it instead
 * of seeding it from the current weights (the reference's EMA callback restores its averages the same way,
 * mode/callbacks/ema.py:150-160). */
int mode_optimizer_ema_mark_seeded(mode_engine_t* e);
/* Sums of squares of n spans (offset, numel; host array) of the flat gradient buffer -> out_dev[n], two launches,
 * deterministic. Replaces the per-parameter `.grad.norm().item()` loop of MoDEAgent.on_before_zero_grad
 * (mode_agent.py:304-359: ~5 host synchronisations per parameter) by one device->host copy of n floats. */
int mode_grad_segment_sumsq(mode_engine_t* e, const int64_t* seg_host, int n, float* out_dev, void* stream);
/* Makes `stream` wait (cudaStreamWaitEvent, no host sync) until the most recent mode_train_step has finished writing the
 * gradients of block `layer` (its backward runs last-to-first), or all gradients when layer == -1. This is what lets a
 * data-parallel caller all-reduce layer l's sections on a side stream while layers l-1..0 are still in backward — the
 * role of DDP's bucketed reducer hooks in the reference (mode/training_calvin.py:97, DDPStrategy). */
int mode_train_wait_grads(mode_engine_t* e, int layer, void* stream);

/* Fused samplers: the whole sigma loop of a k-diffusion sampler over GCDenoiser as ONE CUDA-graph launch (the update is
 * the epilogue of the output-head kernel). `sampler`: MODE_SAMPLER_DDIM = sample_ddim (gc_sampling.py:922-951, the
 * reference default), MODE_SAMPLER_EULER = sample_euler with s_churn = 0 (:164-211), MODE_SAMPLER_DPMPP_2M =
 * sample_dpmpp_2m (:699-734). Arguments as mode_sample_ddim. */
enum { MODE_SAMPLER_DDIM = 0, MODE_SAMPLER_EULER = 1, MODE_SAMPLER_DPMPP_2M = 2 };
int mode_sample(mode_engine_t* e, int sampler, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                const float* sigmas_host, int n_plus_1, int B, void* stream);
/* sample_ddim (gc_sampling.py:922-951) over GCDenoiser: x_inout_dev (B, action_seq_len, action_dim) holds the initial
 * noise (randn * sigma_max, drawn by the caller as in mode_agent.py:756) and receives the denoised actions.
 * sigmas_host: n_plus_1 values, the last one normally 0 (get_sigmas_exponential, gc_sampling.py:35-38).
 * The whole n-step loop is one CUDA graph launch; observation/goal tokens are embedded once per call. */
int mode_sample_ddim(mode_engine_t* e, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                     const float* sigmas_host, int n_plus_1, int B, void* stream);

/* Same as mode_sample_ddim with HOST buffers: copies inputs host->device, runs, copies the result back into
 * x_inout_host and synchronises the stream. This is the end-to-end call a non-PyTorch host would make. */
int mode_sample_ddim_host(mode_engine_t* e, const float* state_host, const float* goal_host, float* x_inout_host,
                          const float* sigmas_host, int n_plus_1, int B, void* stream);

/* Sampler programs: the remaining k-diffusion samplers of gc_sampling.py — sample_heun :257, sample_dpm_2 :315,
 * sample_lms :430, sample_dpmpp_2s :956, sample_euler_ancestral :214, sample_dpm_2_ancestral :376,
 * sample_dpmpp_2s_ancestral :874 — as ONE CUDA-graph launch. Each of their updates is a linear combination of the step's
 * base sample X, the probe P handed to a second evaluation, the evaluation's denoised output D, up to four history
 * tensors H and a caller-drawn noise tensor. Evaluation i runs the denoiser at sigma_eval_host[i] on X
 * (reads_probe_host[i] = 0) or P (1); the head kernel's epilogue then applies row i of prog_host (16 floats):
 *   {cX, cP, cD, cH0, cH1, cH2, cH3, cN, hX, hD, slot, dst, 0, 0, 0, 0}:
 *   dst <- cX*X + cP*P + cD*D + sum_j cHj*H[j] + cN*noise_i   (dst = X if prog[11] == 0 else P)
 *   H[slot] <- hX*x_in + hD*D  if slot >= 0                    (x_in = the evaluation's input)
 * noise_dev: (n_evals, B, A, action_dim) fp32 drawn by the caller (RNG stays with the caller, mode_agent.py:756), or
 * NULL. x_inout_dev receives the final X. n_evals <= 64, every sigma > 0. */
int mode_sample_program(mode_engine_t* e, const float* state_dev, const float* goal_dev, float* x_inout_dev,
                        const float* sigma_eval_host, const int32_t* reads_probe_host, const float* prog_host,
                        const float* noise_dev, int n_evals, int B, void* stream);

/* NoiseBlockMoE.forward(x, c) (modedit.py:530-595), eval mode, for one layer: x_dev/out_dev (B, T, d) with
 * T = 2 + n_state_tokens + action_seq_len, c_dev (B, d). */
int mode_block_forward(mode_engine_t* e, int layer, const float* x_dev, const float* c_dev, float* out_dev, int B,
                       void* stream);

/* Routing of the most recent evaluation, layer `layer` (RouterCond.forward outputs, modedit.py:312-318):
 * idx_host (B, top_k) int32 in torch.topk order, w_host (B, top_k) renormalised probabilities, probs_host (B, E)
 * clamped softmax; any may be NULL. Synchronises the device. */
int mode_get_routing(mode_engine_t* e, int layer, int B, int32_t* idx_host, float* w_host, float* probs_host);
/* The same for evaluation `step` (0-based) of the most recent fused sampler call (mode_sample*): the whole sigma schedule
 * is routed ahead of the loop, one table slot per network evaluation (gc_sampling.py:945 feeds sigmas[i] to every
 * sample, so each (step, layer) has one routing decision). step = -1: the most recent evaluation (as above). */
int mode_get_routing_at(mode_engine_t* e, int step, int layer, int B, int32_t* idx_host, float* w_host, float* probs_host);

/* NoiseBlockMoE.get_expert_usage / total_tokens_processed / reset_expert_usage (modedit.py:597-605). Synchronises. */
int mode_get_expert_usage(mode_engine_t* e, int layer, int64_t* usage_host /* [num_experts] */,
                          int64_t* total_tokens_host /* [1] */);
int mode_reset_expert_usage(mode_engine_t* e);

/* Kernels of this library enqueued by the most recent forward/denoise/sample/block call (graph nodes count). */
int64_t mode_last_launch_count(const mode_engine_t* e);

/* Measurement entry: runs mode_denoise `reps` times with a CUDA-event pair around every kernel launch and returns the
 * mean device time per evaluation (ms) and the launch count of each kernel class, MODE_PROF_CLASSES entries each:
 * 0 router+plan, 1 embed, 2 QKV GEMM, 3 attention, 4 c_proj GEMM, 5 ln_2+permute, 6 expert up GEMM (SwiGLU),
 * 7 expert down GEMM, 8 combine(+ln_1), 9 head, 10 obs/goal embedding. Synchronises the stream. */
#define MODE_PROF_CLASSES 11
int mode_profile_eval(mode_engine_t* e, const float* state_dev, const float* goal_dev, const float* actions_dev,
                      const float* sigma_dev, int sigma_stride, float* out_dev, int B, int reps, void* stream,
                      float* ms_host, int32_t* launches_host);

/* Unit-test entry for the tcgen05 GEMM: out = epilogue(A[M,K] @ W[N,K]^T). a_dev/w_dev bf16, bias_dev/resid_dev fp32
 * (may be NULL), epilogue = 0 bias->bf16, 1 resid+acc->f32, 2 swiglu->bf16 (W/bias already interleaved per 256 rows),
 * 3 plain->bf16, 4 plain->f32. M arbitrary, N % 256 == 0, K % 64 == 0. Returns after enqueueing. */
int mode_debug_gemm(const void* a_dev, const void* w_dev, const float* bias_dev, const float* resid_dev,
                    void* out_dev, int M, int N, int K, int epilogue, void* stream);
/* Unit-test entry for the weight-gradient GEMM: out[N_out, K_out] (fp32) = dY[rows, N_out]^T @ X[rows, K_out]
 * (bf16 operands read MN-major straight from the row-major activations). rows % 64 == 0, N_out % 128 == 0,
 * K_out % 256 == 0; swiglu_half > 0 un-interleaves packed SwiGLU rows on store. */
int mode_debug_wgrad(const void* dy_dev, const void* x_dev, float* out_dev, int rows, int n_out, int k_out,
                     int swiglu_half, void* stream);
/* Unit-test entry for the attention kernel: qkv_dev bf16 (B*T, 3*H*Dh), out_dev bf16 (B*T, H*Dh). */
int mode_debug_attention(const void* qkv_dev, const float* q_gain_dev, const float* k_gain_dev, void* out_dev,
                         int B, int T, int H, int Dh, float eps, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * FiLM-ResNet-50 token producer: the reference's FiLMResNet50Policy
 * (mode/models/perceptual_encoders/pretrained_resnets.py:25-60), which MoDEAgent.embed_visual_obs (mode_agent.py:548-567)
 * calls once per camera to turn images into the `state_images` tokens of the denoiser. Inference only: BatchNorm uses its
 * running statistics and is folded into the convolution weights; every convolution runs as a tcgen05 GEMM over NHWC bf16
 * activations with bias, shortcut add, ReLU and the stage's FiLM modulation in the epilogue.
 * mode_resnet_set_weight takes the module's state_dict keys ("resnet.conv1.weight", "resnet.bn1.running_var",
 * "resnet.layer2.0.downsample.0.weight", "film3.gamma.weight", ...; "*.num_batches_tracked" is accepted and ignored),
 * fp32, reference shapes, host or device. mode_resnet_forward: images_dev (N, 3, H, W) fp32 with the H, W given at create
 * time, cond_dev (N, cond_dim) fp32 (the language goal embedding), out_dev (N, 2048) fp32 pooled features. */
typedef struct mode_resnet mode_resnet_t;
int mode_resnet_create(int cond_dim, int max_images, int height, int width, mode_resnet_t** out);
void mode_resnet_destroy(mode_resnet_t* r);
int mode_resnet_set_weight(mode_resnet_t* r, const char* name, const float* data, int is_device, const int64_t* shape, int ndim);
int mode_resnet_finalize(mode_resnet_t* r, void* stream);
int mode_resnet_forward(mode_resnet_t* r, const float* images_dev, const float* cond_dev, float* out_dev, int N, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MODE_ENGINE_H_ */
