// ffmp_kernels.cuh — kernel argument blocks and launch entry points (internal to libffmp_b200.so).
#pragma once
#include <cuda.h>

#include "ffmp_common.cuh"

namespace ffmp {

// Work items: item -> (env, episode).  slot_mode=1: planes live at [(episode % S) * N + env] (env
// batch with S resident scenario slots); slot_mode=0: planes live at [item] (stateless operators).
struct ScenarioArgs {
    const uint32_t *env_idx;    // dev u32[count] or null (env = item)
    const uint32_t *env_gid;    // dev u32[count] or null (gid = env_id_base + env)
    const uint32_t *episode;    // dev u32[count] or null (use episode_const)
    const uint32_t *count_ptr;  // dev count or null (use `count`)
    int count;
    uint32_t episode_const;
    int G, goal_mode, block_shift, slot_mode, S, N;
    uint32_t p_thresh, env_id_base;
    uint64_t seed;
    uint8_t *occ;
    uint32_t *scen;
};

struct FlowArgs {
    const uint32_t *env_idx;    // dev u32[count] or null
    const uint32_t *episode;    // dev u32[count] or null (use episode_const; slot_mode only)
    const uint32_t *count_ptr;  // dev count or null
    int count;
    uint32_t episode_const;
    int G, slot_mode, S, N;
    const uint8_t *occ;
    const int32_t *goal_cells;  // dev i32[count][2] (slot_mode=0)
    const uint32_t *scen;       // scenario records (slot_mode=1: goal cell read from here)
    int32_t *cost;              // may be null
    uint8_t *flow;
    uint32_t *hi_scratch;       // per-CTA spill area for cost bit-planes >= 8
    uint32_t *ticket;           // dev u32: CTA completion ticket or null; the last CTA re-arms ticket / work / count_reset
    uint32_t *work;             // dev u32 (zero before the launch): dynamic grid hand-out counter, or null for striding
    uint32_t *count_reset;      // dev u32 reset to 0 by the last CTA (the regen list counter) or null
    // in-kernel scenario generation (generate = 1, slot_mode only): SPEC.md §3 parameters and the record output
    int generate, goal_mode, block_shift;
    uint32_t p_thresh, env_id_base;
    uint64_t seed;
    uint32_t *scen_out;
    uint32_t neg1, one;         // 0xFFFFFFFF and 1 (set by the launcher): opaque IMAD multipliers, see flow_field.cu
    uint32_t *host_done;        // optional word in mapped host memory: the last CTA publishes host_done_value there, so the host
    uint32_t host_done_value;   //   can tell without a CUDA call that this (background regeneration) launch has completed
    int all_slots;              // 1 (generate, 96 < G <= 128): the launch covers every scenario slot of N envs — item -> env = item % N,
                                //    episode = episode_const + item / N, count = N * slots (a full reset in ONE launch)
    int latency;                // 1: the caller waits for this launch (a join's flush, a reset of a few envs): 96 < G <= 128 uses the
                                //    four-warps-per-grid kernel
};

struct StepArgs {
    int N, G, W, K, S, max_steps;
    float dt;
    int mode;                   // 0 step, 1 masked reset, 2 begin-all (after full regeneration)
    int slot_new;               // ring slot of the newest frame (older = slot_new-1)
    int write_older;            // 1: every env also (re)writes the older frame
    const int64_t *actions;     // dev i64[N] (mode 0)
    const uint8_t *mask;        // dev u8[N] (mode 1)
    const uint8_t *flow;        // [S][N][G][G]
    const uint32_t *scen;       // [S][N][8]
    uint32_t *state;            // [N][16]
    uint8_t *frames;            // [N][K][W][W]
    float *rel_goal, *velocity, *reward, *term_rel_goal, *term_velocity, *fin_return;
    uint8_t *done, *flags;
    int32_t *fin_length;
    uint32_t *regen_env, *regen_episode, *regen_count;  // regeneration request list
    uint32_t *obs_order;        // [N][8] crop orders, dynamics_kernel -> observe_kernel
    int32_t *term_order;        // [N][8] or null: (ci, cj, pi, pj, plane) of the envs that finished, step kernel -> terminal_obs_kernel
    uint8_t *term_frames;       // [N][2][W][W] or null: terminal local_map of the envs that finished in this step
    uint32_t *error_word;
    unsigned long long *trace;  // optional [N][8] per-CTA timestamps of the tick kernel (diagnostics, FFMP_TRACE=1) or null
};

struct RewarderArgs {
    int n, W;
    const int32_t *local_map;   // map-based collision (ffmp.py:85-105) or null
    const double *scan;         // LiDAR ranges f64[n][scan_len], NaN = None (ffmp.py:108-117) or null
    int scan_len;
    const uint8_t *given_flags; // explicit bit0 collision / bit1 goal (ffmp.py:130 signature) or null
    const float *rel_goal;
    const uint8_t *is_first;
    float *d_first, *reward;
    uint8_t *done, *flags;
};

struct FeedArgs {
    int N, K, W, slot_new, bf16;
    float scale;
    const uint8_t *frames;      // [N][K][W][W]
    void *out;                  // f32 or bf16 [N][2][W][W]
};

struct ScanArgs {
    int n, G, beams;
    float range_max;
    int flow_mode;              // 1: map is a flow image (255 = occupied), 0: an occupancy plane (non-zero = occupied)
    const uint8_t *map;         // u8 planes [..][G][G]
    const float *pose;          // f32[n][3] (operator: plane = item) or null (env: pose and plane from the state records)
    const uint32_t *state;      // env: [N][16]
    int S, N;                   // env: plane = (episode % S) * N + env
    float *scan;                // f32[n][beams]
    uint8_t *hit;               // u8[n] is_collision2 of the beam list, or null
};

constexpr int FEED_MAX_WORLD = 16;

struct FeedTargets {
    uint32_t *word[FEED_MAX_WORLD];   // word[r]: a flag / ack word in rank r's memory (peer mapped), or null
};

struct FeedPushArgs {
    int N, K, W, slot_new, ndst;
    const uint8_t *frames;            // [N][K][W][W] observation ring of this rank
    const float *rel_goal, *velocity, *reward;
    const uint8_t *done;
    uint8_t *dst[FEED_MAX_WORLD];     // this rank's slot in the gather buffer of destination d (peer mapped or local)
    uint32_t *flag[FEED_MAX_WORLD];   // flag word of this rank in destination d's memory
    uint32_t *ticket;                 // local CTA completion counter (wraps to 0 by itself)
    uint32_t seq;
};

struct HostExportArgs {
    const void *src;            // dev: packed result block (reward | rel_goal | velocity | done | flags), 16-byte aligned
    void *dst;                  // device-visible address of the caller's pinned host block, 16-byte aligned
    int n16, tail_words;        // 16-byte words, then 0..3 trailing 32-bit words
    uint32_t *flag;             // device-visible address of the library's mapped completion word
    uint32_t value;             // sequence number published when the block has been written
    unsigned long long *stamps; // optional (diagnostics): three globaltimer stamps in mapped memory, or null
};

// launchers (each returns the cudaError_t of the launch)
cudaError_t launch_host_export(const HostExportArgs &a, cudaStream_t st);   // host_io.cu: programmatic dependent of the tick
cudaError_t launch_feed_spin(const uint32_t *words, int stride_words, uint32_t mask, uint32_t want, double timeout_s,
                             uint32_t *error_word, cudaStream_t st);         // feed.cu
cudaError_t launch_feed_signal(const FeedTargets &t, uint32_t mask, uint32_t value, cudaStream_t st);
cudaError_t launch_feed_push(const FeedPushArgs &a, cudaStream_t st);
cudaError_t launch_scan(const ScanArgs &a, cudaStream_t st);                // scan.cu
cudaError_t launch_learner_input(const FeedArgs &a, cudaStream_t st);
cudaError_t launch_scenarios(const ScenarioArgs &a, int grid, cudaStream_t st);
cudaError_t launch_flow_field(const FlowArgs &a, int grid, cudaStream_t st);
// tmap: TMA descriptor of the flow planes [S*N][G][G] (box W x ceil16(W)) or null -> plain-load observe kernel
// fused: run the scalar step inside the TMA observe kernel (one launch per tick); ignored without a tensor map
// act_bytes: host pointer to N <= ACT_PARAM_MAX action bytes (255 = invalid) that travel inside the launch as a by-value kernel
//            parameter (fused TMA kernel only; the caller checks launch_step_takes_action_bytes), or null (a.actions on the device)
constexpr int ACT_PARAM_MAX = 4096;
cudaError_t launch_step(const StepArgs &a, const CUtensorMap *tmap, cudaStream_t st, cudaEvent_t between = nullptr,
                        bool fused = true, bool pdl = false, const uint8_t *act_bytes = nullptr);
cudaError_t launch_rewarder(const RewarderArgs &a, cudaStream_t st);
cudaError_t launch_terminal_obs(const StepArgs &a, cudaStream_t st);      // step.cu: behind a mode-0 step when term_frames is set
int flow_field_max_grid(int G);            // resident CTAs for a full wave (multiple of the SM count)
size_t flow_field_scratch_words(int G);    // hi_scratch words per CTA
bool flow_field_supported(int G);
bool flow_field_takes_all_slots(int G);    // FlowArgs.all_slots (a full reset in one launch) is available for this grid size

}  // namespace ffmp
