/* hirl_b200.h -- C ABI of libhirl_b200.so (sm_100a).
 *
 * The reference (zrc0622/HIRL4UCAV) is pure Python and has no FFI of its own: its "operator
 * interface" for the hot path is the Python surface that hirl/train_all.py touches.  This
 * header is the boundary a maintainer binds (ctypes) so that those
 * Python entry points run on B200.  Every entry point cites the reference code it replaces
 * (paths relative to the reference checkout).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types, never throws.
 *   - return 0 on success, <0 on error (HIRL_E_*); hirl_last_error() gives a message
 *     (thread-local).
 *   - an opaque handle owns all device memory of one agent; calls on one handle must be
 *     serialised by the caller, different handles are independent.
 *   - "host" pointers are ordinary (ideally pinned) host memory and the call is synchronous;
 *     "_dev" entry points take device pointers plus a cudaStream_t (passed as void*) and
 *     are asynchronous.
 *   - all numerics are IEEE fp32 (env thresholds/arithmetic in fp64 like the reference's
 *     Python floats); there is NO CPU fallback anywhere in this library.
 */
#ifndef HIRL_B200_H
#define HIRL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIRL_ABI_VERSION 1

enum {
    HIRL_OK = 0,
    HIRL_E_INVALID = -1,   /* bad argument (the Python layer raises ValueError)          */
    HIRL_E_CUDA = -2,      /* CUDA runtime failure (message carries cudaGetErrorString)   */
    HIRL_E_UNSUPPORTED = -3, /* shape/feature outside the fused path (NotImplementedError) */
    HIRL_E_NOMEM = -4
};

/* fixed network geometry of the reference driver (hirl/train_all.py:200-203) */
#define HIRL_STATE_DIM 13
#define HIRL_ACTION_DIM 4
#define HIRL_H1 256
#define HIRL_H2 512
#define HIRL_ROW_WORDS 32 /* one replay row: s[13] a[4] s'[13] r done -> 128 B           */
#define HIRL_BC_WORDS 20  /* one expert (s,a) row: s[13] a[4] pad[3]  ->  80 B            */
#define HIRL_OUT_SCALARS 8

/* network selector, state_dict order inside each (hirl/agents/HIRL.py:172-180) */
enum { HIRL_NET_ACTOR = 0, HIRL_NET_TARGET_ACTOR = 1, HIRL_NET_CRITIC = 2,
       HIRL_NET_TARGET_CRITIC = 3, HIRL_NET_BC_ACTOR = 4 };
#define HIRL_ACTOR_PARAMS 138756  /* flat state_dict length of Actor  (HIRL.py:105-126)   */
#define HIRL_CRITIC_PARAMS 276482 /* flat state_dict length of Critic (HIRL.py:19-53)     */

enum { HIRL_BUF_ONLINE = 0, HIRL_BUF_EXPERT = 1 };

/* activation of the hidden layers: HIRL uses relu (HIRL.py:61-66,131-132), the TD3/BC
 * files use leaky_relu(0.01) (TD3.py:61-66, BC.py:129-130). */
enum { HIRL_ACT_RELU = 0, HIRL_ACT_LEAKY_RELU = 1 };

typedef struct hirl_agent hirl_agent_t;

typedef struct hirl_config {
    int32_t device;          /* CUDA device ordinal                                          */
    int32_t batch;           /* batchSize (HIRL.py:158); any value >= 1                      */
    int32_t online_capacity; /* bufferSize (HIRL.py:187; train_all.py:190 uses 1e5)          */
    int32_t expert_capacity; /* len(expert_states)+10 (HIRL.py:190)                          */
    int32_t bc_rows;         /* len(expert_states): rows of the BC table (HIRL.py:245-249)   */
    int32_t max_steps;       /* max learn() steps fused into one launch (>=1)                */
    int32_t activation;      /* HIRL_ACT_*                                                   */
    int32_t replicas;        /* independent agents sharing this handle (ensemble), >=1       */
    float gamma;             /* HIRL.py:156                                                  */
    float tau;               /* HIRL.py:155                                                  */
    float noise_clamp;       /* TD3LearningNoiseClamp, 0.5 (HIRL.py:163)                     */
    float loss_lambda;       /* 10000 (HIRL.py:182)                                          */
} hirl_config_t;

/* One learn() call (HIRL.py:221-334).  The host owns the three RNG streams and the
 * schedule (which calls train the actor, when Polyak fires, Adam bias corrections), exactly
 * the state the reference keeps in Python attributes; the device executes. */
typedef struct hirl_step {
    int32_t n_online;        /* rows drawn from the online ring  (batch - expert_num_now)    */
    int32_t n_expert;        /* rows drawn from the expert ring  (HIRL.py:224-227)           */
    int32_t do_actor;        /* actorTrainable for this call (HIRL.py:291,332)               */
    int32_t do_polyak;       /* update_count % 3 == 0 after this call (HIRL.py:327-330)      */
    float bc_weight_now;     /* 100 selects the soft Q-filter weight (HIRL.py:299)           */
    float bc_warm_up;        /* bc_warm_up_weight (HIRL.py:306)                              */
    float force_soft_weight; /* >=0: override the soft count (parity tests only), else <0    */
    float noise[4];          /* the call's N(0,0.2) draw, unclamped (HIRL.py:265)            */
    float critic_neg_step;   /* -lr/(1-beta1^t) of the critic's Adam (torch single-tensor)   */
    float critic_bc2_sqrt;   /* sqrt(1-beta2^t)                                              */
    float actor_neg_step;
    float actor_bc2_sqrt;
    int32_t mode;            /* 0 HIRL (HIRL.py), 1 TD3 (TD3.py:201-260: no BC terms),
                                2 BC (BC.py:160-185: MSE on an expert batch, actor Adam only)  */
} hirl_step_t;

/* ---- lifetime ------------------------------------------------------------------------- */
int hirl_abi_version(void);
const char *hirl_last_error(void);
/* replaces Agent.__init__ device-side state (HIRL.py:149-190) */
int hirl_agent_create(const hirl_config_t *cfg, hirl_agent_t **out);
int hirl_agent_destroy(hirl_agent_t *h);
int hirl_agent_sm_count(hirl_agent_t *h);
/* which update kernel this handle launches: 0 stage kernel, 1 tensor-core cluster kernel (single relu agent, batch <= 256),
 * 2 FP32 cluster kernel (HIRL_CLUSTER=1), 3 ensemble kernel (replicas > 1) */
int hirl_agent_backend(hirl_agent_t *h);

/* ---- data-parallel large-batch mode --------------------------------------
 * No reference counterpart (the reference is one process); it sits exactly at the reference's two
 * optimizer boundaries: critic_optimizer.step() (HIRL.py:286-288) and actor_optimizer.step()
 * (HIRL.py:323-325), plus the batch mean of the soft Q-filter (HIRL.py:303).  One process per GPU,
 * each a full replica of the agent holding cfg.batch = global_batch / world rows of every call's batch;
 * the per-shard gradients (276 482 + 138 756 fp32), the filter count and the loss scalars are summed
 * with NCCL all-reduces over NVLink between stage ranges of the update kernel, so every rank applies the
 * same Adam step.  Loss means run over global_batch.  The caller cuts the index streams it drew once
 * into shards (hirl4ucav_b200/parallel.py: shard_step).
 *   hirl_nccl_unique_id: rank 0 fills a 128-byte ncclUniqueId, the caller broadcasts it by any side channel;
 *   hirl_allreduce_init: collective; creates this handle's own communicator from the id;
 *   hirl_allreduce_attach: use an existing ncclComm_t instead (not destroyed with the handle). */
int hirl_nccl_unique_id(void *id128);
/* FUSED variant (preferred; the NCCL variant above is the fallback where CUDA IPC between the ranks is unavailable): the
 * all-reduces happen INSIDE the update kernel over NVLink peer memory, a call stays one launch.  hirl_dp_fused_export
 * allocates this rank's exchange block and writes its 64-byte cudaIpcMemHandle_t; the caller gathers all ranks' handles
 * ([world][64], rank order) and passes them to hirl_dp_fused_attach (world <= 8: one node).  Every rank adds the ranks'
 * contributions in rank order, so the summed gradients and therefore the weights are bit-identical everywhere. */
int hirl_dp_fused_export(hirl_agent_t *h, int rank, int world, int global_batch, void *handle64);
int hirl_dp_fused_attach(hirl_agent_t *h, const void *handles, int world);
int hirl_allreduce_init(hirl_agent_t *h, const void *id128, int rank, int world, int global_batch);
int hirl_allreduce_attach(hirl_agent_t *h, void *nccl_comm, int rank, int world, int global_batch);

/* ---- parameters: state_dict import/export (HIRL.py:99-103,142-146,336-350) -------------
 * flat fp32 vector = concatenation of the module's parameters in state_dict order, each
 * tensor row-major (out,in); length HIRL_ACTOR_PARAMS or HIRL_CRITIC_PARAMS.  host ptrs.  */
int hirl_params_import(hirl_agent_t *h, int replica, int net, const float *src, size_t n);
int hirl_params_export(hirl_agent_t *h, int replica, int net, float *dst, size_t n);
/* Adam moments of actor / critic (torch.optim.Adam state 'exp_avg', 'exp_avg_sq') */
int hirl_adam_import(hirl_agent_t *h, int replica, int net, const float *m, const float *v, size_t n);
int hirl_adam_export(hirl_agent_t *h, int replica, int net, float *m, float *v, size_t n);
/* last gradients of actor / critic in the same flat order (parity tests) */
int hirl_grads_export(hirl_agent_t *h, int replica, int net, float *dst, size_t n);

/* ---- replay (hirl/utils/buffer.py:20-36; HIRL.py:214-215, 245-246) --------------------- */
/* writes n_rows consecutive ring slots starting at first_slot (wraps at capacity) */
int hirl_store_rows(hirl_agent_t *h, int replica, int buffer, int first_slot, int n_rows, const float *rows32);
int hirl_bc_table_set(hirl_agent_t *h, const float *states13, const float *actions4, int n_rows);
/* buffer.py:38-48 + HIRL.py:229-233: rows of the online ring first, then the expert ring */
int hirl_gather(hirl_agent_t *h, int replica, const int32_t *idx_online, int n_online,
                const int32_t *idx_expert, int n_expert, float *out_rows32);
/* HIRL.py:249-251 */
int hirl_gather_bc(hirl_agent_t *h, const int32_t *idx_bc, int n, float *out_states13, float *out_actions4);

/* ---- the fused update (HIRL.py:221-334 x k_steps, one persistent launch) ---------------
 * idx_* are [k_steps][batch] int32 (only the first n_online / n_expert of a row are read);
 * out_scalars is [k_steps][8]: critic_loss, actor_loss, bc_loss, rl_loss, bc_fire_loss,
 * bc_weight (HIRL.py:334; the last five keep their previous value on non-actor calls),
 * soft_count, reserved.  With replicas>1 every array has a leading [replicas] dimension. */
int hirl_learn(hirl_agent_t *h, int k_steps, const hirl_step_t *steps, const int32_t *idx_online,
               const int32_t *idx_expert, const int32_t *idx_bc, float *out_scalars);
/* hirl_learn split in two: enqueue (H2D of the host streams, launch, D2H of out_scalars into the caller's
 * pinned buffer) and wait.  The host may draw the NEXT call's indices between the two. */
int hirl_learn_async(hirl_agent_t *h, int k_steps, const hirl_step_t *steps, const int32_t *idx_online,
                     const int32_t *idx_expert, const int32_t *idx_bc, float *out_scalars);
int hirl_learn_wait(hirl_agent_t *h);
/* same with DEVICE pointers, asynchronous on `stream` (a cudaStream_t) */
int hirl_learn_dev(hirl_agent_t *h, int k_steps, const hirl_step_t *steps_dev, const int32_t *idx_online_dev,
                   const int32_t *idx_expert_dev, const int32_t *idx_bc_dev, float *out_scalars_dev, void *stream);
/* number of kernels the last hirl_learn* call launched (bench bookkeeping) */
int hirl_last_launch_count(hirl_agent_t *h);

/* ---- acting (HIRL.py:192-212): actions = net(states), n rows, no noise (host adds it) -- */
int hirl_act(hirl_agent_t *h, int replica, int net, const float *states13, int n, float *actions4);

/* ---- device-resident collection (batched acting + on-device store) ----------
 * A vectorised rollout keeps observations, actions and rewards on the GPU: hirl_env_step_dev
 * produces obs/reward/flags, hirl_act_dev the next actions, hirl_store_transitions_dev appends
 * the transitions to the replay ring -- no per-step H2D/D2H (reference: one chooseAction +
 * one store per env step, HIRL.py:192-215, train_all.py:343-351).
 * noise4_dev (may be NULL): noise_rows==1 -> one 4-vector for all rows, noise_rows==n -> one per
 * row; actions = clamp(tanh(...) + noise, -1, 1) exactly as HIRL.py:196-198. */
int hirl_act_dev(hirl_agent_t *h, int replica, int net, const float *states13_dev, int n, float *actions4_dev,
                 const float *noise4_dev, int noise_rows, void *stream);
/* ring append of n transitions at first_slot (wraps at capacity): row = s[13] a[4] s2[13] r done,
 * done = flags bit 0 (the layout hirl_env_step_dev writes); buffer.py:20-36 for n rows at once. */
int hirl_store_transitions_dev(hirl_agent_t *h, int replica, int buffer, int first_slot, int n, const float *s13_dev,
                               const float *a4_dev, const float *s2_13_dev, const float *reward_dev,
                               const int32_t *flags_dev, void *stream);
/* the handle's own cudaStream_t (updates run on it): lets a caller order its streams against it */
int hirl_agent_stream(hirl_agent_t *h, void **stream_out);

/* ---- host-side index draws (buffer.py:45; HIRL.py:249, BC.py:167) -----------------------
 * Exact replicas of random.sample(range(n), k) (CPython Random.sample) and of
 * np.random.choice(n, k, replace=False) (legacy RandomState: permutation(n)[:k]) on an MT19937
 * state {key[624], pos}: same indices, same final generator state.  `st` is either numpy's own
 * live state (bit_generator.ctypes.state_address has this layout) or a caller-owned copy used for
 * speculative pre-draws.  Pure host code. */
typedef struct hirl_mt19937 { uint32_t key[624]; int32_t pos; } hirl_mt19937_t;
int hirl_host_sample_indices(hirl_mt19937_t *st, int32_t n, int32_t k, int32_t *out);
int hirl_host_choice_noreplace(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t *out);
/* the same draw on a library-owned worker thread (one job at a time): `st` must be a private copy that stays alive,
 * together with `out`, until hirl_host_async_wait() returns.  HIRL_E_UNSUPPORTED if a job is already pending. */
int hirl_host_choice_noreplace_async(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t *out);
/* `count` consecutive draws on the same state (the cloning batches of `count` fused learn() calls): out[count][size] */
int hirl_host_choice_many_async(hirl_mt19937_t *st, int32_t population, int32_t size, int32_t count, int32_t *out);
int hirl_host_async_wait(void);
/* all index draws of k_calls consecutive learn() calls for n_jobs independent generator pairs (ensemble replicas), on
 * n_threads worker threads.  Per job and call, in the reference's order: random.sample(range(online_len), n_online[k]),
 * random.sample(range(expert_len), n_expert[k]) if n_expert[k] > 0 (both on *py), np.random.choice(bc_population, batch,
 * replace=False) on *np (skipped when idx_bc is NULL).  Outputs are [n_jobs][k_calls][batch] int32 (leading entries used). */
typedef struct hirl_draw_job { hirl_mt19937_t *py; hirl_mt19937_t *np; int32_t online_len; int32_t expert_len; } hirl_draw_job_t;
int hirl_host_draw_many(const hirl_draw_job_t *jobs, int32_t n_jobs, int32_t k_calls, const int32_t *n_online, const int32_t *n_expert,
                        int32_t bc_population, int32_t batch, int32_t *idx_online, int32_t *idx_expert, int32_t *idx_bc, int32_t n_threads);

/* ---- test hooks ------------------------------------------------------------------------ */
/* stop the next hirl_learn after this many stages of its first step (0 = run everything) */
int hirl_debug_stop_after(hirl_agent_t *h, int stage);
/* copy a scratch buffer to the host; returns its element count in *n_out when dst==NULL */
int hirl_debug_export(hirl_agent_t *h, int buffer_id, int pass, float *dst, size_t n, size_t *n_out);

/* per-stage SM-clock stamps of the first 12 steps of a launch: [12][17 stages][160 CTAs][2]
 * (work finished, barrier released); enable!=0 switches recording on, dst!=NULL copies out */
int hirl_debug_profile(hirl_agent_t *h, int enable, long long *dst, size_t n);

/* ---- Harfang env math (hirl/environments/HarfangEnv_GYM.py) ----------------------------
 * raw state = 16 words: p_ally[3] e_ally[3] p_opp[3] e_opp[3] target_angle_deg health
 * action_fire flags(u32: b0 locked_prev b1 locked_now b2 missile_prev b3 missile_now
 * b4 fire_success).  Outputs: obs[13] fp32 (GYM.py:193-268), reward fp32 (GYM.py:101-137),
 * flags int32: b0 done b1 episode_success (GYM.py:160-169) b2 fire_success b3 fired
 * (the policy's rule a[3] > 0, GYM.py:150) b4 expert_fire (the expert's rule obs[7] > 0 and obs[8] > 0 on
 * the new observation, data/straight_line/ai_data_col.py:63-64) byte1 step_success int8.  counters (may be NULL): int64[3] running totals
 * fired-with-missile, successful fires, dones (Infinite env, GYM.py:519-527). */
int hirl_env_step_dev(const float *raw16_dev, int64_t n, float *obs13_dev, float *reward_dev,
                      int32_t *flags_dev, unsigned long long *counters3_dev, void *stream);
int hirl_env_step(int device, const float *raw16, int64_t n, float *obs13, float *reward, int32_t *flags,
                  unsigned long long *counters3);
/* offline functions on normalised observations (GYM.py:299-336; train_all.py:295-297):
 * reward fp32, step_success int8, done uint8 of next_state */
int hirl_env_offline_dev(const float *s13_dev, const float *a4_dev, const float *s2_13_dev, int64_t n,
                         float *reward_dev, int8_t *success_dev, uint8_t *done_dev, void *stream);
int hirl_env_offline(int device, const float *s13, const float *a4, const float *s2_13, int64_t n,
                     float *reward, int8_t *success, uint8_t *done);
/* the same on float64 inputs: the reference evaluates get_reward/get_termination on the float64 expert CSV
 * (train_all.py:295-297 on read_data's arrays), and a threshold such as next_state[12] <= 0.1 (GYM.py:334)
 * differs between 0.1 and float32(0.1) */
int hirl_env_offline_f64(int device, const double *s13, const double *a4, const double *s2_13, int64_t n,
                         float *reward, int8_t *success, uint8_t *done);
/* expert-buffer construction (train_all.py:287-306): from the (n,13)/(n,4) expert table builds
 * the stored transitions; out_rows32 holds up to n-1 rows, *n_rows_out how many were kept. */
int hirl_expert_scan(int device, const float *states13, const float *actions4, int n, float *out_rows32,
                     int8_t *out_success, int *n_rows_out);
/* float64 expert table (what read_data returns): thresholds in float64, stored rows cast to fp32 */
int hirl_expert_scan_f64(int device, const double *states13, const double *actions4, int n, float *out_rows32,
                         int8_t *out_success, int *n_rows_out);

/* ---- measurement helper -----------------------------------------------------------------
 * Register-resident FFMA loop on every SM, timed with CUDA events: the FP32-FMA roofline
 * denominator for the batch-256 update (MEASURED_PEAKS.json has HBM and bf16 only). */
int hirl_measure_fp32_peak(int device, double *tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* HIRL_B200_H */
