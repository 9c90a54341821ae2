/*
 * ffmp_b200.h — C-ABI of the B200-native gym_ffmp hot path (libffmp_b200.so).
 *
 * Plain C: pointers, sizes and PODs only (no torch / C++ types).  Every pointer marked "dev" is
 * device memory allocated by the CALLER (e.g. a torch CUDA tensor); the library never frees it.
 * Every `stream` argument is a cudaStream_t passed as void* (torch's current stream); the library
 * never synchronises the host unless the function says so.  One host thread per GPU (one process
 * per GPU under torchrun).  All functions return 0 on success or a negative ffmp_status; the text of
 * the last failure on the calling thread is available from ffmp_last_error().  Nothing throws.
 * There is NO CPU fallback: ffmp_create fails with FFMP_ERR_DEVICE if the device is not sm_100.
 *
 * Reference interfaces replaced (paths under /root/reference/):
 *   - gym.make('FFMP-v0') -> FFMP object            src/train.py:456, src/gym_ffmp/__init__.py:3-6
 *   - FFMP.rewarder / rewarder2 (per-tick reward)   src/gym_ffmp/envs/ffmp.py:167-188, call site src/train.py:577
 *   - RobotAction.commander (action table)          src/gym_ffmp/envs/robot/config.py:25-58, src/train.py:668-669
 *   - ROSNode geometry helpers                      src/train.py:157-188
 *   - make_temporal_maps / obs assembly             src/train.py:474-486, 535-557
 *   - external ROS producers the reference only subscribes to (flow image /bev/temporal_bev_image,
 *     /odom physics, /start_goal_points scenario generator): src/train.py:84-90,116-143
 * The frozen semantics of every entry point are in SPEC.md.
 */
#ifndef FFMP_B200_H
#define FFMP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FFMP_ABI_VERSION 1u
#define FFMP_COST_INF 0x7FFFFFFF

typedef enum ffmp_status {
    FFMP_OK = 0,
    FFMP_ERR_ARG = -1,      /* bad argument / unsupported configuration        */
    FFMP_ERR_DEVICE = -2,   /* no CUDA device, or device is not sm_100          */
    FFMP_ERR_CUDA = -3,     /* a CUDA runtime call failed (see ffmp_last_error) */
    FFMP_ERR_STATE = -4,    /* call order violated (e.g. step before bind/reset) */
    FFMP_ERR_ABI = -5       /* cfg.abi_version mismatch                          */
} ffmp_status;

typedef struct ffmp_handle ffmp_handle;

/* Batched-environment configuration (defaults = the reference's constants, SPEC.md §1). */
typedef struct ffmp_cfg {
    uint32_t abi_version;  /* FFMP_ABI_VERSION */
    int32_t device;        /* CUDA ordinal */
    int32_t num_envs;      /* N: environments on this GPU */
    int32_t grid;          /* G: global grid side, G%4==0, 16..1024            (reference map: 100, ffmp.py:15) */
    int32_t window;        /* W: local map side, W%4==0, <= 256                (ffmp.py:15: 100) */
    int32_t ring;          /* K >= 2 observation frame slots per env (SPEC.md §8) */
    int32_t slots;         /* S in [2,16] resident scenario slots per env (pre-generated episodes) */
    int32_t max_steps;     /* truncation                                       (train.py:60: 200) */
    int32_t goal_mode;     /* 0 goal re-sampled every episode, 1 static goal at (G-8,G-8) */
    int32_t block_shift;   /* obstacle block = 2^block_shift cells */
    uint32_t p_thresh;     /* floor(p_occ * 2^32) */
    uint32_t env_id_base;  /* global id of local env 0 (multi-GPU sharding) */
    uint64_t seed;
    float dt;              /* integration step [s] */
    uint32_t regen_batch;  /* ticks per background regeneration launch: 0 = library default (max(1, min(4, (S-1)/5))), else
                              1..min(8, S-1).  The episode ends of m consecutive ticks share one list and one launch, and the
                              S-1 ticks of slack are split over floor((S-1)/m) lists; the step results do not depend on it */
} ffmp_cfg;

/* Byte sizes of every caller-allocated buffer for a given cfg (ffmp_query_sizes). */
typedef struct ffmp_sizes {
    size_t cost;           /* i32 [S][N][G][G]  integration field                 */
    size_t flow;           /* u8  [S][N][G][G]  flow image (255 occ, else dir*28) */
    size_t scen;           /* u32 [S][N][8]     scenario records                  */
    size_t state;          /* u32 [N][16]       env state records                 */
    size_t frames;         /* u8  [N][K][W][W]  observation frame ring            */
    size_t vec2;           /* f32 [N][2]        rel_goal, velocity, term_*        */
    size_t vec1;           /* 4B  [N]           reward, fin_return, fin_length    */
    size_t bytes1;         /* u8  [N]           done, flags                       */
    size_t workspace;      /* library scratch (regen lists, high cost bit-planes) */
} ffmp_sizes;

/* Scenario record layout (u32 words): 0 x0(f32) 1 y0 2 yaw0 3 gx 4 gy 5 gi(i32) 6 gj 7 hash key.
 * State record layout (u32 words): 0 x 1 y 2 yaw 3 gx 4 gy 5 d_first 6 ep_return (all f32)
 *                                  7 steps(i32) 8 episode(u32) 9..15 reserved. */
typedef struct ffmp_buffers {
    int32_t *cost;             /* dev */
    uint8_t *flow;             /* dev */
    uint32_t *scen;            /* dev */
    uint32_t *state;           /* dev */
    uint8_t *frames;           /* dev: observation local_map ring */
    float *rel_goal;           /* dev f32[N][2]: [dist, bearing]             (train.py:174-180) */
    float *velocity;           /* dev f32[N][2]: [|dxy|, wrap(dyaw)]         (train.py:182-188) */
    float *reward;             /* dev f32[N]                                 (ffmp.py:130-157)  */
    uint8_t *done;             /* dev u8[N]: collision | goal | truncated    (ffmp.py:160-164, train.py:607-608) */
    uint8_t *flags;            /* dev u8[N]: bit0 collision bit1 goal bit2 truncated */
    float *term_rel_goal;      /* dev f32[N][2]: pre-reset values of the step just taken */
    float *term_velocity;      /* dev f32[N][2] */
    float *fin_return;         /* dev f32[N]: episode return, valid where done */
    int32_t *fin_length;       /* dev i32[N]: episode length, valid where done */
    void *workspace;           /* dev, ffmp_sizes.workspace bytes, zero-initialised by the caller */
} ffmp_buffers;

const char *ffmp_last_error(void);
uint32_t ffmp_abi_version(void);

int ffmp_query_sizes(const ffmp_cfg *cfg, ffmp_sizes *out);
int ffmp_create(const ffmp_cfg *cfg, ffmp_handle **out);
/* Alignment (checked, FFMP_ERR_ARG otherwise): cost / flow / scen / state / frames / workspace 16 bytes; rel_goal / velocity /
 * term_rel_goal / term_velocity 8 bytes (written as float2); reward / fin_return / fin_length 4 bytes.  A packed output block
 * [reward f32 N | rel_goal f32 2N | velocity f32 2N | done u8 N | flags u8 N] therefore needs an even N.                 */
int ffmp_bind(ffmp_handle *h, const ffmp_buffers *bufs);
int ffmp_destroy(ffmp_handle *h);

/* reset: mask_dev == NULL -> every env restarts from episode 0 (all S scenario slots are generated,
 * flow fields included, on `stream`).  Otherwise masked envs (mask != 0) abandon their episode and
 * start their next one (same path as an auto-reset).  Observation outputs are written.            */
int ffmp_reset(ffmp_handle *h, const uint8_t *mask_dev, void *stream);

/* step: actions_dev = i64[N] ids in [0,28) (SPEC.md §1).  Writes every output buffer, auto-resets
 * finished envs, and queues the regeneration of their consumed scenario slot (scenario generation fused
 * into the flow-field kernel) on internal background streams.  Replaces the per-tick body of train.py:531-682.                             */
int ffmp_step(ffmp_handle *h, const int64_t *actions_dev, void *stream);

/* T back-to-back steps with actions_dev = i64[T][N]; identical to T ffmp_step calls. */
int ffmp_rollout(ffmp_handle *h, const int64_t *actions_dev, int32_t T, void *stream);

/* The same T steps replayed as ONE CUDA graph launch.  The first call for a given (actions_dev, T, ring phase, regeneration
 * list phase) captures the T step kernels and their background regeneration launches (stream capture, side streams forked
 * and joined), instantiates the graph and launches it; later calls with the same key only launch it (up to 16 graphs are
 * kept, least recently used first out).  actions_dev must stay valid and is re-read by every replay.  A graph ends with
 * the join of its regenerations, so `stream` needs no ffmp_join afterwards.  Results are identical to ffmp_rollout.    */
int ffmp_rollout_graphed(ffmp_handle *h, const int64_t *actions_dev, int32_t T, void *stream);

/* Host-buffer step (the reference-facing call with HOST memory): actions_host i64[N] in, reward f32[N], done u8[N],
 * flags u8[N], rel_goal f32[N][2] and velocity f32[N][2] out (any out pointer may be NULL); returns when the results are
 * in the host buffers.  The local_map observation stays on the device.
 *   - Fast path: the five bound device outputs are adjacent in memory in the order reward | rel_goal | velocity | done |
 *     flags, the five host destinations are adjacent in the same order (one 22*N-byte block each, 16-byte aligned, N even)
 *     and the host block is pinned (cudaHostAlloc / cudaHostRegister).  Then a small kernel queued behind the step writes
 *     the host block directly and publishes a completion word in mapped memory which the call spins on: no device-to-host
 *     copy engine, no stream synchronisation.  Up to 4096 envs the actions travel INSIDE the step kernel's launch: the call
 *     narrows them to one byte per env (255 for anything outside [0, 28), which the kernel flags in the error word like an
 *     out-of-range int64) and passes the block as a by-value kernel parameter — no copy engine on the way in either
 *     (FFMP_ACT_PARAM=0: off).  Above 4096 envs they go in with one cudaMemcpyAsync (8 N bytes).
 *   - Otherwise: cudaMemcpyAsync in, step, cudaMemcpyAsync out (one copy when both blocks are packed), stream sync.
 * FFMP_HOST_IO in the environment at ffmp_create selects the path: 0 copy engines both ways, 1 (default) mapped results,
 * 2 mapped results and actions read in place over PCIe by the step kernel (faster on some hosts, slower on others:
 * profiles/r01g_e2e_paths.txt, r01h_e2e_paths.txt).
 * ffmp_step_host = ffmp_step_host_async + ffmp_step_host_wait (gym.vector's step_async / step_wait): the host may
 * work between the two; exactly one wait per async.                                                             */
int ffmp_step_host(ffmp_handle *h, const int64_t *actions_host, float *reward_host, uint8_t *done_host,
                   uint8_t *flags_host, float *rel_goal_host, float *velocity_host, void *stream);
int ffmp_step_host_async(ffmp_handle *h, const int64_t *actions_host, float *reward_host, uint8_t *done_host,
                         uint8_t *flags_host, float *rel_goal_host, float *velocity_host, void *stream);
int ffmp_step_host_wait(ffmp_handle *h);

/* Per-kernel device timing (CUDA events recorded by the library on the launching streams).  enable != 0: start
 * recording the next (at most 256) ticks.  enable == 0: stop, synchronise the recorded events and return the
 * averages in milliseconds: tick_ms = the step kernel(s) on the caller's stream, regen_ms = the background
 * scenario-regeneration (flow-field) launch of the tick on its side stream.                             */
int ffmp_timing(ffmp_handle *h, int32_t enable, float *tick_ms, float *regen_ms, int32_t *ticks);

/* Number of kernels this handle has launched since ffmp_create (step, reset and regeneration launches). */
int ffmp_launch_count(const ffmp_handle *h, uint64_t *out);

/* Diagnostics: with FFMP_TRACE=1 in the environment at ffmp_create, the tick kernel records eight timestamps per env
 * (u64[N][8]: globaltimer at start, clock<<8|smid, clock after kinematics, before / after the window wait, at the
 * start of the drain, at the end, globaltimer at the end) of its latest launch.  Copies them out; synchronises `stream`. */
int ffmp_debug_trace(ffmp_handle *h, uint64_t *out_host, void *stream);

/* Index of the newest frame slot p (1 <= p <= K-1): the observation is frames[:, p-1 : p+1]. */
int ffmp_obs_slot(const ffmp_handle *h, int32_t *newest_slot);

/* Resume support: set the newest frame slot after the caller restored the bound buffers from a checkpoint (a full
 * ffmp_reset must have run on this handle first).                                                                 */
int ffmp_set_obs_slot(ffmp_handle *h, int32_t newest_slot);

/* The current observation as the learner's input tensor (replaces make_temporal_maps + `.float()` + `.to(device)`,
 * src/train.py:474-486, 539-545): out = f32 (dtype 0) or bf16 (dtype 1) [N][2][W][W], oldest frame first, pixel value
 * times `scale` (1.0 = the reference's plain float cast).  out_dev: caller-allocated device memory, 16-byte aligned.   */
int ffmp_learner_input(ffmp_handle *h, void *out_dev, int32_t dtype, float scale, void *stream);

/* LiDAR scan synthesis (SPEC.md §9) on every env's current scenario and pose: scan f32[N][beams] ranges in metres
 * (+inf = no return within range_max, 0 = robot cell occupied), hit u8[N] = FFMP.is_collision2 of the beam list (may be
 * NULL).  Replaces the /scan LaserScan the reference fed to rewarder2 (src/train.py:87,144-150,577; ffmp.py:108-117). */
int ffmp_scan(ffmp_handle *h, int32_t beams, float range_max, float *scan_dev, uint8_t *hit_dev, void *stream);

/* Make `stream` wait (device-side, no host sync) for all queued background regeneration. */
/* Terminal observations (SPEC.md §7; the reference's last tick of an episode, /root/reference/src/train.py:611-664):
 * term_frames_dev = caller-allocated u8[N][2][W][W] or NULL (off, the default).  When set, every ffmp_step / ffmp_rollout /
 * ffmp_step_host tick is followed by one small kernel that writes, for the envs whose `done` is set, the terminal local_map
 * [previous frame, crop at the terminal pose on the finished episode's flow image]; rows of other envs are left untouched. */
int ffmp_set_terminal_obs(ffmp_handle *h, uint8_t *term_frames_dev);

int ffmp_join(ffmp_handle *h, void *stream);

/* Device error word accumulated by the kernels (bit0: action id out of range). Synchronises `stream`. */
int ffmp_error_word(ffmp_handle *h, uint32_t *out, void *stream);

/* ---- learner feed over NVLink peer memory (replaces pack + NCCL all-gather of the transition block) ----
 * One ffmp_feed per rank (one process per GPU).  The library allocates a gather buffer [2][world][slot] on `device`; the
 * ranks exchange its 64-byte CUDA IPC handle out of band (torch.distributed all_gather_object) and map each other's buffer
 * with ffmp_feed_connect.  Every step, on every rank (SPMD, same order):
 *   ffmp_feed_push(env, feed, dest_mask, ...)   one kernel reads the rank's transition block where the step left it (two
 *       newest ring frames, relative_goal, velocity, reward, done) and stores it in the packed layout
 *       [maps u8 N*2*W*W | rel_goal f32 N*2 | velocity f32 N*2 | reward f32 N | done u8 N] into slot `rank` of every
 *       destination rank's buffer (bit r of dest_mask) over NVLink, then publishes the push sequence number there;
 *   ffmp_feed_wait(feed, src_mask, ...)         stream-ordered (a spin kernel, no host sync) until the listed sources'
 *       blocks of the current sequence number have arrived; *buffer_dev = [world][slot_stride] of this sequence number;
 *   ffmp_feed_release(feed, src_mask, ...)      after the consumer's reads (stream order): returns the buffer to the
 *       producers.  Two buffers alternate, so a producer is at most one step ahead of its slowest consumer.
 * block_bytes must equal N * (2 W^2 + 21).  A wait that exceeds timeout_s sets bit 1 of the feed's error word.        */
typedef struct ffmp_feed ffmp_feed;
#define FFMP_IPC_HANDLE_BYTES 64
int ffmp_feed_create(int32_t device, int32_t world, int32_t rank, size_t block_bytes, ffmp_feed **out);
int ffmp_feed_handle(ffmp_feed *f, uint8_t *handle_out);
int ffmp_feed_connect(ffmp_feed *f, int32_t peer_rank, const uint8_t *handle);
int ffmp_feed_info(const ffmp_feed *f, void **base_dev, size_t *slot_stride, size_t *buffer_stride, uint32_t *seq);
int ffmp_feed_push(ffmp_handle *h, ffmp_feed *f, uint32_t dest_mask, double timeout_s, void *stream);
int ffmp_feed_wait(ffmp_feed *f, uint32_t src_mask, double timeout_s, void **buffer_dev, void *stream);
int ffmp_feed_release(ffmp_feed *f, uint32_t src_mask, void *stream);
int ffmp_feed_error(ffmp_feed *f, uint32_t *out, void *stream);
int ffmp_feed_destroy(ffmp_feed *f);

/* The current transition block of the env, packed [maps u8 N*2*W*W | rel_goal f32 N*2 | velocity f32 N*2 | reward f32 N |
 * done u8 N] (N * (2 W^2 + 21) bytes) into caller memory on this device: the storage step of a device replay ring
 * (ReplayMemory.push, src/train.py:212-228) and the single-GPU form of ffmp_feed_push.  dst_dev: 16-byte aligned.  */
int ffmp_pack_transitions(ffmp_handle *h, void *dst_dev, void *stream);

/* ---- stateless operators (what the reference's external ROS nodes computed) ------------------- */

/* Scenario generator (SPEC.md §3): for item n, env id env_gid_dev[n] (u32) and episode episode_dev[n]
 * (u32) -> occ u8[n][G][G], scen u32[n][8].  Replaces the external episode_manager node
 * (train.py:86-90,128-132).                                                                        */
int ffmp_op_scenarios(int32_t device, int32_t n, int32_t G, uint32_t p_thresh, int32_t goal_mode, int32_t block_shift,
                      uint64_t seed, const uint32_t *env_gid_dev, const uint32_t *episode_dev, uint8_t *occ_dev,
                      uint32_t *scen_dev, void *stream);

/* Flow field (SPEC.md §4,§5): occ u8[n][G][G] + goal cells i32[n][2] -> cost i32[n][G][G] (may be NULL),
 * flow u8[n][G][G].  workspace_dev: ffmp_op_flow_field_workspace(n,G) bytes.  Replaces the external
 * /bev/* flow-image node (train.py:84,116-121).                                                    */
size_t ffmp_op_flow_field_workspace(int32_t n, int32_t G);
int ffmp_op_flow_field(int32_t device, int32_t n, int32_t G, const uint8_t *occ_dev, const int32_t *goal_cells_dev,
                       int32_t *cost_dev, uint8_t *flow_dev, void *workspace_dev, void *stream);

/* Stateless LiDAR scan synthesis (SPEC.md §9): map u8[n][G][G] (flow_mode 1: flow image, 255 = occupied; 0: occupancy
 * plane, non-zero = occupied), pose f32[n][3] (x, y, yaw) -> scan f32[n][beams], hit u8[n] (may be NULL).       */
int ffmp_op_scan(int32_t device, int32_t n, int32_t G, const uint8_t *map_dev, int32_t flow_mode, const float *pose_dev,
                 int32_t beams, float range_max, float *scan_dev, uint8_t *hit_dev, void *stream);

/* Batched FFMP.rewarder (ffmp.py:167-176) on caller-supplied ego-centred local maps:
 * local_map i32[n][W][W] (>0 = occupied, robot at (W/2,W/2)), rel_goal f32[n][2], is_first u8[n],
 * d_first f32[n] (in/out latch, per item instead of the reference's module global)
 * -> reward f32[n], done u8[n], flags u8[n] (bit0 collision, bit1 goal).                           */
int ffmp_op_rewarder(int32_t device, int32_t n, int32_t W, const int32_t *local_map_dev, const float *rel_goal_dev,
                     const uint8_t *is_first_dev, float *d_first_dev, float *reward_dev, uint8_t *done_dev,
                     uint8_t *flags_dev, void *stream);

/* Batched FFMP.rewarder2 (ffmp.py:179-188, the variant train.py:577 calls): scan f64[n][scan_len] LiDAR
 * ranges, NaN = None; collision iff a non-zero range is < 0.13 (ffmp.py:108-117).                  */
int ffmp_op_rewarder2(int32_t device, int32_t n, int32_t scan_len, const double *scan_dev, const float *rel_goal_dev,
                      const uint8_t *is_first_dev, float *d_first_dev, float *reward_dev, uint8_t *done_dev,
                      uint8_t *flags_dev, void *stream);

/* Batched FFMP.reward_calculator (ffmp.py:130-157): collision / goal given explicitly in given_flags
 * (bit0 collision, bit1 goal).                                                                      */
int ffmp_op_reward_calculator(int32_t device, int32_t n, const float *rel_goal_dev, const uint8_t *given_flags_dev,
                              const uint8_t *is_first_dev, float *d_first_dev, float *reward_dev, uint8_t *done_dev,
                              uint8_t *flags_dev, void *stream);

/* Minibatch gather of a replay ring of packed transition blocks (ReplayMemory.sample + Brain.make_minibatch,
 * /root/reference/src/train.py:224-225, 349-369) in one kernel: index i64 [B][2] = (push number k >= 1, env e); state = push
 * k-1, action / reward / done / next observation = push k.  Observation stacks come out as bf16 NCHW [B][2][W][W]. */
int ffmp_replay_gather(int32_t device, const uint8_t *blocks_dev, size_t stride, int32_t T, int32_t N, int32_t W,
                       const int64_t *actions_dev, const int64_t *index_dev, int32_t B, void *state_m_bf16,
                       void *observe_m_bf16, float *state_g, float *state_v, float *observe_g, float *observe_v,
                       float *reward, uint8_t *done, int64_t *action, void *stream);

/* ---- Q network of the reference trainer -------------------------------------------------------------
 * Network.forward(state_m, state_g, state_v, state_t) of /root/reference/src/train.py:231-303 as tcgen05 implicit-GEMM kernels
 * (csrc/qnet.cu): bf16 operands, fp32 accumulation in TMEM.  Weights are given in the layouts of the module's state_dict
 * (f32, device), in the order conv1, conv2, conv3, conv4, fc1, fc2, fc3, fc4_ea, fc4_ev; ffmp_qnet_load converts them once.
 * state_m: [B][2][100][100] NCHW, dtype 1 = bfloat16 (what ffmp_learner_input writes), 0 = float32; state_g / state_v f32
 * [B][2], state_t f32 [B][1]; q_out f32 [B][28].  scalar_tile != 0 keeps the reference's scalar tile x_gvt_[0][30]
 * (train.py:264-276: the only path by which goal / velocity / dt reach the output); 0 leaves it out. */
typedef struct ffmp_qnet ffmp_qnet;
int ffmp_qnet_create(int32_t device, int32_t max_batch, ffmp_qnet **out);
int ffmp_qnet_load(ffmp_qnet *n, const float *const *weights_dev, const float *const *biases_dev, void *stream);
int ffmp_qnet_forward(ffmp_qnet *n, int32_t batch, const void *state_m_dev, int32_t dtype, const float *state_g_dev,
                      const float *state_v_dev, const float *state_t_dev, int32_t scalar_tile, float *q_out_dev, void *stream);
/* layer 1..6: the six convolution outputs (bf16 NHWC), 7: fc2, 8: fc3 (bf16 [B][512]); out_dev may be NULL (size query) */
int ffmp_qnet_debug_activation(ffmp_qnet *n, int32_t layer, int32_t batch, void *out_dev, size_t *elems_per_sample, void *stream);
int ffmp_qnet_launch_count(const ffmp_qnet *n, uint64_t *out);
int ffmp_qnet_destroy(ffmp_qnet *n);
const char *ffmp_qnet_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* FFMP_B200_H */
