// ffmp_api.cu — the C-ABI of libffmp_b200.so (include/ffmp_b200.h): handle, buffer binding, stream /
// event plumbing for the background scenario regeneration, and the stateless operators.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <new>
#include <string>
#include <vector>

#include "../../include/ffmp_b200.h"
#include "ffmp_kernels.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char *what, cudaError_t ce = cudaSuccess) {
    g_err = what;
    if (ce != cudaSuccess) {
        g_err += ": ";
        g_err += cudaGetErrorString(ce);
    }
    return code;
}

#define CK(call)                                                  \
    do {                                                          \
        cudaError_t ce_ = (call);                                 \
        if (ce_ != cudaSuccess) return fail(FFMP_ERR_CUDA, #call, ce_); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        if (switched) cudaSetDevice(prev);
    }
};

constexpr int MAX_LISTS = 15;
constexpr int REGEN_GRID = 148 * 2;   // CTAs of a background regeneration launch per tick of its group (grids are handed out dynamically)

double now_us() {
    timespec ts;
    clock_gettime(CLOCK_REALTIME, &ts);
    return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3;
}

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int check_device(int device) {
    // cudaGetDeviceProperties costs milliseconds: the verdict is cached per ordinal
    static int verdict[64] = {0};   // 0 unknown, 1 ok
    if (device >= 0 && device < 64 && verdict[device] == 1) return FFMP_OK;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(FFMP_ERR_DEVICE, "no CUDA device available (there is no CPU fallback)");
    }
    if (device < 0 || device >= n) return fail(FFMP_ERR_DEVICE, "device ordinal out of range");
    int major = 0, minor = 0;
    cudaError_t ce = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (ce == cudaSuccess) ce = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, device);
    if (ce != cudaSuccess) return fail(FFMP_ERR_CUDA, "cudaDeviceGetAttribute", ce);
    if (major != 10 || minor != 0)
        return fail(FFMP_ERR_DEVICE, "device is not sm_100 (B200); this library carries sm_100a code only");
    if (device < 64) verdict[device] = 1;
    return FFMP_OK;
}

int check_cfg(const ffmp_cfg *c) {
    if (!c) return fail(FFMP_ERR_ARG, "cfg is null");
    if (c->abi_version != FFMP_ABI_VERSION) return fail(FFMP_ERR_ABI, "cfg.abi_version != FFMP_ABI_VERSION");
    if (c->num_envs <= 0) return fail(FFMP_ERR_ARG, "num_envs must be > 0");
    if (c->grid < 16 || c->grid > 1024 || (c->grid % 4)) return fail(FFMP_ERR_ARG, "grid must be a multiple of 4 in [16,1024]");
    if (!ffmp::flow_field_supported(c->grid)) return fail(FFMP_ERR_ARG, "grid must be <= 128, or a multiple of 32 up to 512");
    if (c->window < 4 || c->window > 256 || (c->window % 4)) return fail(FFMP_ERR_ARG, "window must be a multiple of 4 in [4,256]");
    if (c->ring < 2 || c->ring > 1024) return fail(FFMP_ERR_ARG, "ring must be in [2,1024]");
    if (c->slots < 2 || c->slots > MAX_LISTS + 1) return fail(FFMP_ERR_ARG, "slots must be in [2,16]");
    if (c->max_steps <= 0) return fail(FFMP_ERR_ARG, "max_steps must be > 0");
    if (c->goal_mode != 0 && c->goal_mode != 1) return fail(FFMP_ERR_ARG, "goal_mode must be 0 or 1");
    if (c->block_shift < 0 || c->block_shift > 8) return fail(FFMP_ERR_ARG, "block_shift must be in [0,8]");
    if (!(c->dt > 0.0f)) return fail(FFMP_ERR_ARG, "dt must be > 0");
    if (c->regen_batch > 8 || static_cast<int>(c->regen_batch) > c->slots - 1)
        return fail(FFMP_ERR_ARG, "regen_batch must be 0 (library default) or in [1, min(8, slots - 1)]");
    return FFMP_OK;
}

// Ticks per regeneration launch (ffmp_cfg.regen_batch).  The episode ends of m consecutive ticks share one regeneration
// list and one background launch; between the ticks of a group nothing but the step kernel is queued, so each of them
// starts as a programmatic dependent of the one before.  An env that ends an episode in the first tick of a group may
// need the regenerated slot S-1 ticks later, so the floor((S-1)/m) lists are reused every nlist*m <= S-1 ticks.  The
// default keeps at least five lists (regenerations in flight).
int regen_batch_of(const ffmp_cfg *c) {
    if (c->regen_batch) return static_cast<int>(c->regen_batch);
    const int m = (c->slots - 1) / 5;
    return m < 1 ? 1 : (m > 4 ? 4 : m);
}

// CTAs of a background regeneration launch: about one grid per CTA for the ~100 episode ends per tick of the bench
// workload (a second round would double the launch's latency), bounded by what the kernel can keep resident.
int regen_grid_of(const ffmp_cfg *c) {
    const int maxg = ffmp::flow_field_max_grid(c->grid);
    const int want = REGEN_GRID * (regen_batch_of(c) + 1) / 2;
    const int g = want < maxg ? want : maxg;
    return c->num_envs < g ? c->num_envs : g;
}

struct Workspace {
    size_t error_word, lists, actions, obs_order, hi, total;
    size_t list_stride;  // bytes per regen list block: [count,ticket,pad..256B][env u32 m*N][episode u32 m*N]
};

Workspace workspace_layout(const ffmp_cfg *c) {
    Workspace w{};
    const size_t N = static_cast<size_t>(c->num_envs);
    size_t off = 0;
    w.error_word = off; off += 256;
    const size_t m = static_cast<size_t>(regen_batch_of(c));
    const size_t nlist = static_cast<size_t>(c->slots - 1) / m;
    w.list_stride = align_up(256 + 2 * m * N * sizeof(uint32_t), 256);   // an env can end one episode per tick
    w.lists = off; off += w.list_stride * nlist;
    w.actions = off; off += align_up(N * sizeof(int64_t), 256);
    w.obs_order = off; off += align_up(N * 8 * sizeof(uint32_t), 256);
    w.hi = off;
    off += align_up((static_cast<size_t>(ffmp::flow_field_max_grid(c->grid)) + nlist * regen_grid_of(c)) *
                        ffmp::flow_field_scratch_words(c->grid) * 4, 256);
    w.total = off;
    return w;
}

}  // namespace

struct ffmp_handle {
    ffmp_cfg cfg;
    ffmp_buffers b;
    Workspace ws;
    bool bound = false, ready = false;
    cudaStream_t side[MAX_LISTS];   // one background stream per regeneration list: regenerations overlap
    int nlist = 1;
    int batch = 1;                  // ticks per regeneration list / launch (regen_batch_of)
    bool group_open = false;        // ticks of the current group have run and its regeneration is not launched yet
    cudaEvent_t ev_step[MAX_LISTS], ev_regen[MAX_LISTS];
    bool regen_pending[MAX_LISTS];
    uint32_t regen_seq[MAX_LISTS];  // launches of list l so far; the kernel publishes it in mapped memory (flag_host[16 + l]) when done
    uint64_t step_index = 0;
    int p = 1;  // newest ring slot
    int ff_grid = 0;                // full-batch grid (reset)
    int rg_grid = 0;                // background regeneration grid (few items per tick)
    CUtensorMap tmap;               // flow planes as a 3-D u8 tensor for the TMA observe kernel
    bool use_tma = false;
    bool always_wait = false;       // FFMP_REGEN_WAIT=1: always queue the stream wait on the regeneration's event (see run_tick)
    bool tick_pdl = true;           // step kernels launched with the programmatic-dependent attribute (FFMP_TICK_PDL=0: off)
    bool fused = true;              // one kernel per tick (env FFMP_STEP_FUSED=0 selects dynamics + observe kernels)
    // optional per-kernel timing (ffmp_timing): events [before tick, after tick] on the caller's stream and
    // [before regeneration, after regeneration] on the side stream of the tick
    static constexpr int TIMING_RING = 256;
    bool timing = false;
    int timing_n = 0;
    cudaEvent_t tev[TIMING_RING][4];
    bool tev_regen[TIMING_RING];    // the tick closed its group: events 2 / 3 bracket a regeneration launch
    uint64_t launches = 0;          // kernels launched by this handle (ffmp_launch_count)
    // host-buffer steps (ffmp_step_host*): completion word in mapped pinned memory, written by host_export_kernel
    volatile uint32_t *flag_host = nullptr;
    uint32_t *flag_dev = nullptr;
    uint32_t flag_seq = 0;
    bool act_param = true;          // host-buffer steps of <= ACT_PARAM_MAX envs pass the actions inside the launch (FFMP_ACT_PARAM=0: copy them)
    uint8_t act_bytes[ffmp::ACT_PARAM_MAX];
    int host_io = 1;                // FFMP_HOST_IO: 0 copy engines + stream sync, 1 mapped results (default), 2 mapped results + in-place actions
    int wait_mode = 0;              // what ffmp_step_host_wait has to do: 0 nothing, 1 spin on the flag, 2 synchronise wait_stream
    cudaStream_t wait_stream = nullptr;
    // FFMP_HOST_IO_STATS=1: host / device timeline of the host-buffer steps, printed to stderr by ffmp_destroy
    bool io_stats = false;
    double t_entry = 0, t_launched = 0, t_queued = 0, t_alias = 0, t_copyin = 0, t_tick = 0, t_evq = 0, acc9 = 0;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double acc8 = 0;
    uint64_t acc_n = 0;
    // ffmp_rollout_graphed: instantiated CUDA graphs of T-tick rollouts, keyed by what is baked into their nodes
    struct RolloutGraph {
        const int64_t *actions; int32_t T; int p0; int list0; const uint8_t *term;   // key
        cudaGraphExec_t exec; int p1; uint64_t ticks; uint64_t launches; uint64_t stamp;
    };
    std::vector<RolloutGraph> graphs;
    uint64_t graph_stamp = 0;
    bool capturing = false;
    cudaStream_t cap_stream = nullptr;     // origin stream of the captures
    uint8_t *term_frames = nullptr;        // caller's [N][2][W][W] buffer (ffmp_set_terminal_obs) or null
    int32_t *term_order = nullptr;         // library-owned [N][8] (allocated with the first ffmp_set_terminal_obs)
    unsigned long long *trace = nullptr;   // FFMP_TRACE=1: [N][8] tick-kernel timestamps (library-owned, diagnostics only)

    uint32_t *error_word() const { return reinterpret_cast<uint32_t *>(static_cast<char *>(b.workspace) + ws.error_word); }
    char *list_base(int l) const { return static_cast<char *>(b.workspace) + ws.lists + ws.list_stride * l; }
    uint32_t *list_count(int l) const { return reinterpret_cast<uint32_t *>(list_base(l)); }
    uint32_t *list_ticket(int l) const { return reinterpret_cast<uint32_t *>(list_base(l)) + 1; }
    uint32_t *list_work(int l) const { return reinterpret_cast<uint32_t *>(list_base(l)) + 2; }
    uint32_t *reset_ticket() const { return error_word() + 4; }      // full-reset launches (same 256-byte header block)
    uint32_t *reset_work() const { return error_word() + 5; }
    uint32_t *list_env(int l) const { return reinterpret_cast<uint32_t *>(list_base(l) + 256); }
    uint32_t *list_episode(int l) const { return list_env(l) + static_cast<size_t>(batch) * cfg.num_envs; }
    int64_t *actions() const { return reinterpret_cast<int64_t *>(static_cast<char *>(b.workspace) + ws.actions); }
    uint32_t *obs_order() const { return reinterpret_cast<uint32_t *>(static_cast<char *>(b.workspace) + ws.obs_order); }
    uint32_t *hi_scratch() const { return reinterpret_cast<uint32_t *>(static_cast<char *>(b.workspace) + ws.hi); }
};

namespace {

ffmp::FlowArgs flow_args(const ffmp_handle *h) {
    ffmp::FlowArgs a{};
    const ffmp_cfg &c = h->cfg;
    a.G = c.grid; a.slot_mode = 1; a.S = c.slots; a.N = c.num_envs;
    a.scen = h->b.scen; a.cost = h->b.cost; a.flow = h->b.flow;
    // scenario generation is fused into the flow-field kernel (no occupancy plane round trip)
    a.generate = 1; a.goal_mode = c.goal_mode; a.block_shift = c.block_shift; a.p_thresh = c.p_thresh;
    a.env_id_base = c.env_id_base; a.seed = c.seed; a.scen_out = h->b.scen;
    a.hi_scratch = h->hi_scratch();
    return a;
}

ffmp::StepArgs step_args(const ffmp_handle *h) {
    ffmp::StepArgs a{};
    const ffmp_cfg &c = h->cfg;
    const ffmp_buffers &b = h->b;
    a.N = c.num_envs; a.G = c.grid; a.W = c.window; a.K = c.ring; a.S = c.slots; a.max_steps = c.max_steps; a.dt = c.dt;
    a.flow = b.flow; a.scen = b.scen; a.state = b.state; a.frames = b.frames;
    a.rel_goal = b.rel_goal; a.velocity = b.velocity; a.reward = b.reward;
    a.term_rel_goal = b.term_rel_goal; a.term_velocity = b.term_velocity; a.fin_return = b.fin_return;
    a.done = b.done; a.flags = b.flags; a.fin_length = b.fin_length;
    a.error_word = h->error_word();
    a.obs_order = h->obs_order();
    a.trace = h->trace;
    a.term_frames = h->term_frames;
    a.term_order = h->term_frames ? h->term_order : nullptr;
    return a;
}

// Queue the background regeneration of the current group's list (the episode ends of its ticks so far) behind the last
// tick on `st`, and round the tick counter up to the next group.
int launch_regen(ffmp_handle *h, cudaStream_t st, cudaEvent_t *tev, bool flush = false) {
    const int l = static_cast<int>((h->step_index / static_cast<uint64_t>(h->batch)) % static_cast<uint64_t>(h->nlist));
    CK(cudaEventRecord(h->ev_step[l], st));
    CK(cudaStreamWaitEvent(h->side[l], h->ev_step[l], 0));
    if (tev) CK(cudaEventRecord(tev[2], h->side[l]));
    ffmp::FlowArgs fa = flow_args(h);
    fa.env_idx = h->list_env(l); fa.episode = h->list_episode(l); fa.count_ptr = h->list_count(l);
    fa.ticket = h->list_ticket(l); fa.work = h->list_work(l); fa.count_reset = h->list_count(l);
    fa.hi_scratch = h->hi_scratch() + (static_cast<size_t>(h->ff_grid) + static_cast<size_t>(l) * h->rg_grid) * ffmp::flow_field_scratch_words(h->cfg.grid);
    fa.host_done = h->flag_dev + 16 + l;
    fa.host_done_value = ++h->regen_seq[l];
    fa.latency = flush ? 1 : 0;
    CK(ffmp::launch_flow_field(fa, h->rg_grid, h->side[l]));
    if (tev) CK(cudaEventRecord(tev[3], h->side[l]));
    h->launches += 1;
    CK(cudaEventRecord(h->ev_regen[l], h->side[l]));
    h->regen_pending[l] = true;
    h->group_open = false;
    h->step_index = (h->step_index / static_cast<uint64_t>(h->batch) + 1) * static_cast<uint64_t>(h->batch);
    return FFMP_OK;
}

// One env-step-like call (mode 0 step, mode 1 masked reset) with the background regeneration queued.
int run_tick(ffmp_handle *h, int mode, const int64_t *actions, const uint8_t *mask, cudaStream_t st,
             const ffmp::HostExportArgs *host_export = nullptr, const uint8_t *act_bytes = nullptr) {
    const uint64_t m = static_cast<uint64_t>(h->batch);
    const int l = static_cast<int>((h->step_index / m) % static_cast<uint64_t>(h->nlist));
    const bool first = h->step_index % m == 0, last = h->step_index % m == m - 1;
    if (first && h->regen_pending[l]) {
        // the regeneration that last used list `l` (nlist groups ago) must be complete: it re-armed the
        // list and filled the scenario slot an env may switch to in this tick.  If the host already sees it complete
        // (the usual case when the caller synchronises every step) no device-side wait is queued in front of the tick.
        // The regeneration kernel's last CTA publishes its launch number in mapped host memory: one plain load tells whether
        // the launch has completed (a cudaEventQuery costs 1.7 us of host time in front of the step kernel's launch).
        // The skipped wait relies on this chain in the regeneration kernel: every CTA's stores, __threadfence, the ticket
        // atomic; the last CTA's __threadfence_system, then the flag store; the host's load of the flag precedes the launch
        // of the tick, whose loads are issued after kernel start (L1 invalidated).  FFMP_REGEN_WAIT=1 always queues the
        // stream wait instead (the parity suite runs in both modes: test_rollout_with_regeneration_wait_forced).
        if (h->capturing || h->always_wait || h->flag_host[16 + l] != h->regen_seq[l]) CK(cudaStreamWaitEvent(st, h->ev_regen[l], 0));
        h->regen_pending[l] = false;
    }
    if (h->io_stats) h->t_evq = now_us();
    ffmp::StepArgs a = step_args(h);
    a.mode = mode; a.actions = actions; a.mask = mask;
    if (mode == 0) {
        if (h->cfg.ring == 2) { a.slot_new = 1; a.write_older = 1; }
        else if (h->p + 1 < h->cfg.ring) { h->p += 1; a.slot_new = h->p; a.write_older = 0; }
        else { h->p = 1; a.slot_new = 1; a.write_older = 1; }
    } else {
        a.slot_new = h->p; a.write_older = 0;
    }
    a.regen_env = h->list_env(l); a.regen_episode = h->list_episode(l); a.regen_count = h->list_count(l);
    const bool one_kernel = h->use_tma && h->fused;
    cudaEvent_t *tev = (h->timing && h->timing_n < ffmp_handle::TIMING_RING) ? h->tev[h->timing_n++] : nullptr;
    if (tev) CK(cudaEventRecord(tev[0], st));
    // the step kernel is always launched as a programmatic dependent of whatever precedes it on the stream (it touches no
    // global memory before its griddepcontrol.wait): behind another step kernel its CTAs are resident when that one drains
    CK(ffmp::launch_step(a, h->use_tma ? &h->tmap : nullptr, st, nullptr, h->fused, h->tick_pdl, act_bytes));
    if (h->io_stats) h->t_tick = now_us();
    if (host_export) {
        // directly behind the step kernel (nothing in between), so that the programmatic dependency pairs the two
        CK(ffmp::launch_host_export(*host_export, st));
        h->launches += 1;
        if (h->io_stats) h->t_launched = now_us();
    }
    if (tev) CK(cudaEventRecord(tev[1], st));
    h->launches += one_kernel ? 1 : 2;
    if (mode == 0 && a.term_frames) {
        // the terminal observations of the envs that finished, before their slot is handed to the regeneration
        CK(ffmp::launch_terminal_obs(a, st));
        h->launches += 1;
    }
    if (!last) {
        h->group_open = true;
        h->step_index += 1;
        if (tev) h->tev_regen[h->timing_n - 1] = false;
        return FFMP_OK;
    }
    if (tev) h->tev_regen[h->timing_n - 1] = true;
    return launch_regen(h, st, tev);
}

void drop_graphs(ffmp_handle *h) {
    for (auto &g : h->graphs) cudaGraphExecDestroy(g.exec);
    h->graphs.clear();
}

// Flush an unfinished group and order `st` after every queued regeneration; later work on `st` needs no further waits.
int join_and_clear(ffmp_handle *h, cudaStream_t st) {
    if (h->group_open) if (int rc = launch_regen(h, st, nullptr, true)) return rc;
    for (int l = 0; l < h->nlist; ++l)
        if (h->regen_pending[l]) { CK(cudaStreamWaitEvent(st, h->ev_regen[l], 0)); h->regen_pending[l] = false; }
    return FFMP_OK;
}

}  // namespace

extern "C" {

const char *ffmp_last_error(void) { return g_err.c_str(); }
uint32_t ffmp_abi_version(void) { return FFMP_ABI_VERSION; }

int ffmp_query_sizes(const ffmp_cfg *cfg, ffmp_sizes *out) {
    if (int rc = check_cfg(cfg)) return rc;
    if (!out) return fail(FFMP_ERR_ARG, "out is null");
    const size_t N = cfg->num_envs, G = cfg->grid, W = cfg->window, K = cfg->ring, S = cfg->slots;
    out->cost = S * N * G * G * sizeof(int32_t);
    out->flow = S * N * G * G;
    out->scen = S * N * ffmp::SC_WORDS * sizeof(uint32_t);
    out->state = N * ffmp::ST_WORDS * sizeof(uint32_t);
    out->frames = N * K * W * W;
    out->vec2 = N * 2 * sizeof(float);
    out->vec1 = N * sizeof(float);
    out->bytes1 = N;
    out->workspace = workspace_layout(cfg).total;
    return FFMP_OK;
}

int ffmp_create(const ffmp_cfg *cfg, ffmp_handle **out) {
    if (!out) return fail(FFMP_ERR_ARG, "out is null");
    *out = nullptr;
    if (int rc = check_cfg(cfg)) return rc;
    if (int rc = check_device(cfg->device)) return rc;
    DeviceGuard guard(cfg->device);
    ffmp_handle *h = new (std::nothrow) ffmp_handle();
    if (!h) return fail(FFMP_ERR_ARG, "out of host memory");
    h->cfg = *cfg;
    if (const char *f = std::getenv("FFMP_STEP_FUSED")) h->fused = std::atoi(f) != 0;
    if (const char *f = std::getenv("FFMP_HOST_IO")) h->host_io = std::atoi(f);
    if (const char *f = std::getenv("FFMP_TICK_PDL")) h->tick_pdl = std::atoi(f) != 0;
    if (const char *f = std::getenv("FFMP_ACT_PARAM")) h->act_param = std::atoi(f) != 0;
    if (const char *f = std::getenv("FFMP_REGEN_WAIT")) h->always_wait = std::atoi(f) != 0;
    if (const char *f = std::getenv("FFMP_HOST_IO_STATS")) h->io_stats = std::atoi(f) != 0;
    h->ws = workspace_layout(cfg);
    h->batch = regen_batch_of(cfg);
    h->nlist = (cfg->slots - 1) / h->batch;
    std::memset(&h->b, 0, sizeof(h->b));
    for (int i = 0; i < MAX_LISTS; ++i) { h->side[i] = nullptr; h->ev_step[i] = nullptr; h->ev_regen[i] = nullptr; h->regen_pending[i] = false; h->regen_seq[i] = 0; }
    std::memset(h->tev, 0, sizeof(h->tev));
    cudaError_t ce = cudaSuccess;
    for (int i = 0; i < h->nlist && ce == cudaSuccess; ++i) {
        ce = cudaStreamCreateWithFlags(&h->side[i], cudaStreamNonBlocking);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_step[i], cudaEventDisableTiming);
        if (ce == cudaSuccess) ce = cudaEventCreateWithFlags(&h->ev_regen[i], cudaEventDisableTiming);
    }
    if (ce == cudaSuccess) {
        void *fh = nullptr, *fd = nullptr;
        ce = cudaHostAlloc(&fh, 256, cudaHostAllocMapped);   // word 0 completion, words 2..7 stamps, words 16..31 regeneration lists
        if (ce == cudaSuccess) {
            std::memset(fh, 0, 256);
            h->flag_host = static_cast<volatile uint32_t *>(fh);
            ce = cudaHostGetDevicePointer(&fd, fh, 0);
            h->flag_dev = static_cast<uint32_t *>(fd);
        }
    }
    if (ce != cudaSuccess) {
        ffmp_destroy(h);
        return fail(FFMP_ERR_CUDA, "stream / event / flag creation", ce);
    }
    if (const char *t = std::getenv("FFMP_TRACE")) {
        if (std::atoi(t) != 0 && cudaMalloc(&h->trace, static_cast<size_t>(cfg->num_envs) * 8 * sizeof(unsigned long long)) != cudaSuccess) {
            cudaGetLastError();
            h->trace = nullptr;
        }
    }
    const int maxg = ffmp::flow_field_max_grid(cfg->grid);
    h->ff_grid = cfg->num_envs < maxg ? cfg->num_envs : maxg;
    h->rg_grid = regen_grid_of(cfg);
    *out = h;
    return FFMP_OK;
}

int ffmp_bind(ffmp_handle *h, const ffmp_buffers *bufs) {
    if (!h || !bufs) return fail(FFMP_ERR_ARG, "null argument");
    const void *ptrs[] = {bufs->cost, bufs->flow, bufs->scen, bufs->state, bufs->frames, bufs->rel_goal,
                          bufs->velocity, bufs->reward, bufs->done, bufs->flags, bufs->term_rel_goal, bufs->term_velocity,
                          bufs->fin_return, bufs->fin_length, bufs->workspace};
    for (const void *p : ptrs)
        if (!p) return fail(FFMP_ERR_ARG, "every ffmp_buffers pointer must be set");
    const uintptr_t aligned[] = {reinterpret_cast<uintptr_t>(bufs->cost),
                                 reinterpret_cast<uintptr_t>(bufs->flow), reinterpret_cast<uintptr_t>(bufs->frames),
                                 reinterpret_cast<uintptr_t>(bufs->state), reinterpret_cast<uintptr_t>(bufs->scen),
                                 reinterpret_cast<uintptr_t>(bufs->workspace)};
    for (uintptr_t p : aligned)
        if (p % 16) return fail(FFMP_ERR_ARG, "plane / frame / state / workspace buffers must be 16-byte aligned");
    // the step kernels write the two-float records with one 8-byte store and the scalars with 4-byte stores
    const void *pairs[] = {bufs->rel_goal, bufs->velocity, bufs->term_rel_goal, bufs->term_velocity};
    for (const void *p : pairs)
        if (reinterpret_cast<uintptr_t>(p) % 8)
            return fail(FFMP_ERR_ARG, "rel_goal / velocity / term_rel_goal / term_velocity must be 8-byte aligned");
    const void *words[] = {bufs->reward, bufs->fin_return, bufs->fin_length};
    for (const void *p : words)
        if (reinterpret_cast<uintptr_t>(p) % 4) return fail(FFMP_ERR_ARG, "reward / fin_return / fin_length must be 4-byte aligned");
    drop_graphs(h);
    h->b = *bufs;
    h->bound = true;
    h->ready = false;
    h->use_tma = false;
    const ffmp_cfg &c = h->cfg;
    if (c.grid % 16 == 0 && ((c.window + 30) & ~15) <= 256) {
        // 3-D tensor map over the flow planes: dims (cols, rows, planes), box = ceil16(W+15) x W x 1, zero fill
        typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                     const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                     CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        DeviceGuard guard(c.device);
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t G = static_cast<cuuint64_t>(c.grid);
            const cuuint64_t dims[3] = {G, G, static_cast<cuuint64_t>(c.slots) * static_cast<cuuint64_t>(c.num_envs)};
            const cuuint64_t strides[2] = {G, G * G};
            const cuuint32_t box[3] = {static_cast<cuuint32_t>((c.window + 30) & ~15), static_cast<cuuint32_t>(c.window), 1};
            const cuuint32_t estr[3] = {1, 1, 1};
            CUresult r = reinterpret_cast<EncodeFn>(fn)(&h->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, bufs->flow, dims, strides, box,
                                                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                                        CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            h->use_tma = r == CUDA_SUCCESS;
        } else {
            cudaGetLastError();
        }
    }
    return FFMP_OK;
}

int ffmp_destroy(ffmp_handle *h) {
    if (!h) return FFMP_OK;
    DeviceGuard guard(h->cfg.device);
    for (int i = 0; i < MAX_LISTS; ++i) {
        if (h->side[i]) {
            cudaStreamSynchronize(h->side[i]);
            cudaStreamDestroy(h->side[i]);
        }
        if (h->ev_step[i]) cudaEventDestroy(h->ev_step[i]);
        if (h->ev_regen[i]) cudaEventDestroy(h->ev_regen[i]);
    }
    if (h->io_stats && h->acc_n) {
        const double n = static_cast<double>(h->acc_n);
        std::fprintf(stderr, "[ffmp host io] host us from entry: pointers classified %.2f | actions copy queued %.2f | regeneration event checked %.2f | step kernel submitted %.2f\n",
                     h->acc[6] / n, h->acc[7] / n, h->acc9 / n, h->acc8 / n);
        std::fprintf(stderr, "[ffmp host io] steps %llu  us/step: submit step+export %.2f | all queued %.2f | results seen %.2f || "
                             "export resident->written %.2f (waiting for the step grid %.2f, copy+fence %.2f)\n",
                     static_cast<unsigned long long>(h->acc_n), h->acc[0] / n, h->acc[1] / n, h->acc[2] / n, h->acc[3] / n,
                     h->acc[4] / n, h->acc[5] / n);
    }
    drop_graphs(h);
    if (h->cap_stream) cudaStreamDestroy(h->cap_stream);
    if (h->trace) cudaFree(h->trace);
    if (h->term_order) cudaFree(h->term_order);
    if (h->flag_host) cudaFreeHost(const_cast<uint32_t *>(h->flag_host));
    if (h->tev[0][0])
        for (int i = 0; i < ffmp_handle::TIMING_RING; ++i)
            for (int j = 0; j < 4; ++j) cudaEventDestroy(h->tev[i][j]);
    delete h;
    return FFMP_OK;
}

int ffmp_reset(ffmp_handle *h, const uint8_t *mask_dev, void *stream) {
    if (!h) return fail(FFMP_ERR_ARG, "handle is null");
    if (!h->bound) return fail(FFMP_ERR_STATE, "ffmp_bind must be called before ffmp_reset");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (mask_dev) {
        if (!h->ready) return fail(FFMP_ERR_STATE, "a full ffmp_reset must precede a masked reset");
        return run_tick(h, 1, nullptr, mask_dev, st);
    }
    // full reset: drain the background stream, regenerate every scenario slot, begin episode 0 everywhere
    for (int l = 0; l < h->nlist; ++l)
        if (h->regen_pending[l]) { CK(cudaStreamWaitEvent(st, h->ev_regen[l], 0)); h->regen_pending[l] = false; }
    const ffmp_cfg &c = h->cfg;
    CK(cudaMemsetAsync(h->b.state, 0, static_cast<size_t>(c.num_envs) * ffmp::ST_WORDS * sizeof(uint32_t), st));
    CK(cudaMemsetAsync(h->b.workspace, 0, h->ws.actions, st));  // error word + regen lists
    if (ffmp::flow_field_takes_all_slots(c.grid)) {
        // the interleaved-layout kernels take every scenario slot in ONE launch (item -> env, slot): a single env's reset is one
        // launch of `slots` grids on the four-warps-per-grid kernel instead of `slots` launches of one grid
        ffmp::FlowArgs fa = flow_args(h);
        const long long total = static_cast<long long>(c.num_envs) * c.slots;
        fa.count = static_cast<int>(total); fa.episode_const = 0; fa.all_slots = 1;
        fa.ticket = h->reset_ticket(); fa.work = h->reset_work();
        const int maxg = ffmp::flow_field_max_grid(c.grid);
        CK(ffmp::launch_flow_field(fa, total < maxg ? static_cast<int>(total) : maxg, st));
        h->launches += 1;
    } else {
        for (int s = 0; s < c.slots; ++s) {
            ffmp::FlowArgs fa = flow_args(h);
            fa.count = c.num_envs; fa.episode_const = static_cast<uint32_t>(s);
            fa.ticket = h->reset_ticket(); fa.work = h->reset_work();
            CK(ffmp::launch_flow_field(fa, h->ff_grid, st));
            h->launches += 1;
        }
    }
    h->p = 1;
    h->step_index = 0;
    h->group_open = false;
    ffmp::StepArgs a = step_args(h);
    a.mode = 2; a.slot_new = 1; a.write_older = 1;
    a.regen_env = h->list_env(0); a.regen_episode = h->list_episode(0); a.regen_count = h->list_count(0);
    CK(ffmp::launch_step(a, h->use_tma ? &h->tmap : nullptr, st, nullptr, h->fused));
    h->launches += (h->use_tma && h->fused) ? 1 : 2;
    h->ready = true;
    return FFMP_OK;
}

int ffmp_step(ffmp_handle *h, const int64_t *actions_dev, void *stream) {
    if (!h || !actions_dev) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_step");
    DeviceGuard guard(h->cfg.device);
    return run_tick(h, 0, actions_dev, nullptr, static_cast<cudaStream_t>(stream));
}

int ffmp_rollout_graphed(ffmp_handle *h, const int64_t *actions_dev, int32_t T, void *stream) {
    if (!h || !actions_dev || T < 0) return fail(FFMP_ERR_ARG, "bad argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_rollout_graphed");
    if (T == 0) return FFMP_OK;
    if (h->timing || h->trace || h->wait_mode) return ffmp_rollout(h, actions_dev, T, stream);   // diagnostics modes: plain launches
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (int rc = join_and_clear(h, st)) return rc;
    const int list0 = static_cast<int>((h->step_index / static_cast<uint64_t>(h->batch)) % static_cast<uint64_t>(h->nlist));
    ffmp_handle::RolloutGraph *g = nullptr;
    for (auto &c : h->graphs)
        if (c.actions == actions_dev && c.T == T && c.p0 == h->p && c.list0 == list0 && c.term == h->term_frames) g = &c;
    if (!g) {
        // capture the T ticks with their regeneration launches (side streams fork from and join back into `st`)
        const uint64_t launches0 = h->launches, ticks0 = h->step_index;
        const int p0 = h->p;
        // (on a stream of the library: the caller's may be the legacy default stream, which cannot be captured)
        if (!h->cap_stream) CK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
        cudaStream_t cs = h->cap_stream;
        CK(cudaStreamBeginCapture(cs, cudaStreamCaptureModeRelaxed));
        h->capturing = true;
        int rc = FFMP_OK;
        for (int32_t t = 0; t < T && rc == FFMP_OK; ++t)
            rc = run_tick(h, 0, actions_dev + static_cast<size_t>(t) * h->cfg.num_envs, nullptr, cs);
        if (rc == FFMP_OK) rc = join_and_clear(h, cs);
        h->capturing = false;
        cudaGraph_t graph = nullptr;
        const cudaError_t ce = cudaStreamEndCapture(cs, &graph);
        if (rc != FFMP_OK || ce != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            cudaGetLastError();
            for (int l = 0; l < h->nlist; ++l) h->regen_pending[l] = false;
            h->ready = false;     // host-side counters advanced without the work having run: a full reset is required
            return rc != FFMP_OK ? rc : fail(FFMP_ERR_CUDA, "stream capture of the rollout", ce);
        }
        cudaGraphExec_t exec = nullptr;
        const cudaError_t ci = cudaGraphInstantiate(&exec, graph, 0);
        cudaGraphDestroy(graph);
        if (ci != cudaSuccess) { h->ready = false; return fail(FFMP_ERR_CUDA, "cudaGraphInstantiate", ci); }
        if (h->graphs.size() >= 16) {     // evict the least recently used graph
            size_t victim = 0;
            for (size_t i = 1; i < h->graphs.size(); ++i) if (h->graphs[i].stamp < h->graphs[victim].stamp) victim = i;
            cudaGraphExecDestroy(h->graphs[victim].exec);
            h->graphs.erase(h->graphs.begin() + static_cast<long>(victim));
        }
        h->graphs.push_back({actions_dev, T, p0, list0, h->term_frames, exec, h->p, h->step_index - ticks0, h->launches - launches0, 0});
        g = &h->graphs.back();
        // the capture advanced the host-side state exactly as the replay below will
        h->p = p0; h->step_index = ticks0; h->launches = launches0;
    }
    CK(cudaGraphLaunch(g->exec, st));
    g->stamp = ++h->graph_stamp;
    h->p = g->p1;
    h->step_index += g->ticks;
    h->launches += g->launches;
    return FFMP_OK;
}

int ffmp_rollout(ffmp_handle *h, const int64_t *actions_dev, int32_t T, void *stream) {
    if (!h || !actions_dev || T < 0) return fail(FFMP_ERR_ARG, "bad argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_rollout");
    DeviceGuard guard(h->cfg.device);
    for (int32_t t = 0; t < T; ++t)
        if (int rc = run_tick(h, 0, actions_dev + static_cast<size_t>(t) * h->cfg.num_envs, nullptr,
                              static_cast<cudaStream_t>(stream)))
            return rc;
    return FFMP_OK;
}

// Device-visible alias of a host pointer, or null when the memory is not pinned / mapped (then the copy engines are used).
static void *mapped_alias(const void *p) {
    if (!p) return nullptr;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

int ffmp_step_host_async(ffmp_handle *h, const int64_t *actions_host, float *reward_host, uint8_t *done_host,
                         uint8_t *flags_host, float *rel_goal_host, float *velocity_host, void *stream) {
    if (!h || !actions_host) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_step_host");
    if (h->wait_mode) return fail(FFMP_ERR_STATE, "ffmp_step_host_wait must be called before the next ffmp_step_host_async");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t N = h->cfg.num_envs;
    if (h->io_stats) h->t_entry = now_us();
    // When the five device outputs are adjacent in memory in the order reward | rel_goal | velocity | done | flags
    // (FFMPVectorEnv allocates them so) and the host destinations are too, they move as ONE 22 N-byte block.
    const ffmp_buffers &b = h->b;
    const char *d0 = reinterpret_cast<const char *>(b.reward);
    char *h0 = reinterpret_cast<char *>(reward_host);
    const bool dev_packed = reinterpret_cast<const char *>(b.rel_goal) == d0 + 4 * N &&
                            reinterpret_cast<const char *>(b.velocity) == d0 + 12 * N &&
                            reinterpret_cast<const char *>(b.done) == d0 + 20 * N &&
                            reinterpret_cast<const char *>(b.flags) == d0 + 21 * N;
    const bool host_packed = reward_host && reinterpret_cast<char *>(rel_goal_host) == h0 + 4 * N &&
                             reinterpret_cast<char *>(velocity_host) == h0 + 12 * N &&
                             reinterpret_cast<char *>(done_host) == h0 + 20 * N &&
                             reinterpret_cast<char *>(flags_host) == h0 + 21 * N;
    const bool packed = dev_packed && host_packed;

    // mapped path: the export kernel writes the caller's pinned block and the completion word; no device-to-host copy, no sync
    void *out_alias = nullptr;
    if (h->host_io >= 1 && packed && (22 * N) % 4 == 0 && reinterpret_cast<uintptr_t>(d0) % 16 == 0 &&
        reinterpret_cast<uintptr_t>(h0) % 16 == 0)
        out_alias = mapped_alias(h0);
    const int64_t *actions_dev = nullptr;
    if (out_alias && h->host_io >= 2) actions_dev = static_cast<const int64_t *>(mapped_alias(actions_host));
    if (h->io_stats) h->t_alias = now_us();
    const uint8_t *act_bytes = nullptr;
    if (!actions_dev && h->act_param && h->use_tma && h->fused && !h->trace && N <= static_cast<size_t>(ffmp::ACT_PARAM_MAX)) {
        // up to ACT_PARAM_MAX envs: the actions ride in the step kernel's launch as one byte per env (255 = out of range, which
        // the kernel reports exactly like an out-of-range int64); no copy engine sits between the caller's buffer and the kernel
        uint8_t *p = h->act_bytes;
        for (size_t i = 0; i < N; ++i) {
            const uint64_t v = static_cast<uint64_t>(actions_host[i]);
            p[i] = v < 28 ? static_cast<uint8_t>(v) : static_cast<uint8_t>(255);
        }
        act_bytes = p;
    }
    if (!actions_dev && !act_bytes) {
        CK(cudaMemcpyAsync(h->actions(), actions_host, N * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        actions_dev = h->actions();
    }
    if (h->io_stats) h->t_copyin = now_us();
    if (out_alias) {
        ffmp::HostExportArgs ea{};
        ea.src = d0; ea.dst = out_alias;
        ea.n16 = static_cast<int>(22 * N / 16);
        ea.tail_words = static_cast<int>((22 * N % 16) / 4);
        ea.flag = h->flag_dev;
        ea.value = ++h->flag_seq;
        if (h->io_stats) ea.stamps = reinterpret_cast<unsigned long long *>(h->flag_dev + 2);
        if (int rc = run_tick(h, 0, actions_dev, nullptr, st, &ea, act_bytes)) return rc;
        h->wait_mode = 1;
        h->wait_stream = st;
        if (h->io_stats) h->t_queued = now_us();
        return FFMP_OK;
    }
    if (int rc = run_tick(h, 0, actions_dev, nullptr, st, nullptr, act_bytes)) return rc;
    if (packed) {
        CK(cudaMemcpyAsync(h0, d0, 22 * N, cudaMemcpyDeviceToHost, st));
    } else {
        if (reward_host) CK(cudaMemcpyAsync(reward_host, b.reward, N * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (done_host) CK(cudaMemcpyAsync(done_host, b.done, N, cudaMemcpyDeviceToHost, st));
        if (flags_host) CK(cudaMemcpyAsync(flags_host, b.flags, N, cudaMemcpyDeviceToHost, st));
        if (rel_goal_host) CK(cudaMemcpyAsync(rel_goal_host, b.rel_goal, N * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
        if (velocity_host) CK(cudaMemcpyAsync(velocity_host, b.velocity, N * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    }
    h->wait_mode = 2;
    h->wait_stream = st;
    return FFMP_OK;
}

int ffmp_step_host_wait(ffmp_handle *h) {
    if (!h) return fail(FFMP_ERR_ARG, "handle is null");
    const int mode = h->wait_mode;
    h->wait_mode = 0;
    if (mode == 0) return FFMP_OK;
    if (mode == 2) {
        DeviceGuard guard(h->cfg.device);
        CK(cudaStreamSynchronize(h->wait_stream));
        return FFMP_OK;
    }
    // spin on the mapped completion word; every 2^14 polls ask the stream whether it died or drained without publishing
    const uint32_t want = h->flag_seq;
    for (uint32_t spins = 1;; ++spins) {
        if (*h->flag_host == want) break;
#if defined(__x86_64__) || defined(__i386__)
        __builtin_ia32_pause();
#endif
        if ((spins & 0x3FFF) == 0) {
            DeviceGuard guard(h->cfg.device);
            const cudaError_t q = cudaStreamQuery(h->wait_stream);
            if (q == cudaErrorNotReady) cudaGetLastError();
            if (q == cudaSuccess) {
                if (*h->flag_host == want) break;
                return fail(FFMP_ERR_CUDA, "the step finished without publishing its completion word");
            }
            if (q != cudaErrorNotReady) return fail(FFMP_ERR_CUDA, "ffmp_step_host_wait", q);
        }
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (h->io_stats) {
        // device stamps are globaltimer ns; they are anchored to the host clock at the flag (seen ~1 us after it is written)
        const double t_seen = now_us();
        const volatile unsigned long long *g = reinterpret_cast<const volatile unsigned long long *>(h->flag_host + 2);
        const double g0 = g[0] * 1e-3, g1 = g[1] * 1e-3, g2 = g[2] * 1e-3;
        h->acc[0] += h->t_launched - h->t_entry;   // entry -> step kernel + export kernel submitted
        h->acc[1] += h->t_queued - h->t_entry;     // entry -> all of the tick's plumbing submitted (spin starts)
        h->acc[2] += t_seen - h->t_entry;          // entry -> results seen
        h->acc[3] += g2 - g0;                      // export kernel resident -> block written
        h->acc[4] += g1 - g0;                      // export kernel resident -> step grid complete
        h->acc[5] += g2 - g1;                      // step grid complete -> block written
        h->acc[6] += h->t_alias - h->t_entry;      // entry -> pointer classification done
        h->acc[7] += h->t_copyin - h->t_entry;     // entry -> actions copy queued
        h->acc8 += h->t_tick - h->t_entry;         // entry -> step kernel submitted
        h->acc9 += h->t_evq - h->t_entry;          // entry -> regeneration event checked
        h->acc_n += 1;
    }
    return FFMP_OK;
}

int ffmp_step_host(ffmp_handle *h, const int64_t *actions_host, float *reward_host, uint8_t *done_host,
                   uint8_t *flags_host, float *rel_goal_host, float *velocity_host, void *stream) {
    if (int rc = ffmp_step_host_async(h, actions_host, reward_host, done_host, flags_host, rel_goal_host, velocity_host, stream))
        return rc;
    return ffmp_step_host_wait(h);
}

int ffmp_timing(ffmp_handle *h, int32_t enable, float *tick_ms, float *regen_ms, int32_t *ticks) {
    if (!h) return fail(FFMP_ERR_ARG, "handle is null");
    DeviceGuard guard(h->cfg.device);
    if (enable) {
        if (!h->tev[0][0])
            for (int i = 0; i < ffmp_handle::TIMING_RING; ++i)
                for (int j = 0; j < 4; ++j) CK(cudaEventCreate(&h->tev[i][j]));
        h->timing = true;
        h->timing_n = 0;
        return FFMP_OK;
    }
    h->timing = false;
    double d = 0, o = 0;
    int regens = 0;
    for (int i = 0; i < h->timing_n; ++i) {
        float a = 0, b = 0;
        CK(cudaEventSynchronize(h->tev[i][1]));
        CK(cudaEventElapsedTime(&a, h->tev[i][0], h->tev[i][1]));
        d += a;
        if (h->tev_regen[i]) {
            CK(cudaEventSynchronize(h->tev[i][3]));
            CK(cudaEventElapsedTime(&b, h->tev[i][2], h->tev[i][3]));
            o += b;
            regens += 1;
        }
    }
    const int n = h->timing_n > 0 ? h->timing_n : 1;
    if (tick_ms) *tick_ms = static_cast<float>(d / n);
    if (regen_ms) *regen_ms = static_cast<float>(o / (regens > 0 ? regens : 1));   // average per regeneration launch
    if (ticks) *ticks = h->timing_n;
    h->timing_n = 0;
    return FFMP_OK;
}

int ffmp_debug_trace(ffmp_handle *h, uint64_t *out_host, void *stream) {
    if (!h || !out_host) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->trace) return fail(FFMP_ERR_STATE, "tracing is off (set FFMP_TRACE=1 before ffmp_create)");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(out_host, h->trace, static_cast<size_t>(h->cfg.num_envs) * 8 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return FFMP_OK;
}

int ffmp_set_terminal_obs(ffmp_handle *h, uint8_t *term_frames_dev) {
    if (!h) return fail(FFMP_ERR_ARG, "handle is null");
    if (h->wait_mode) return fail(FFMP_ERR_STATE, "a host-buffer step is pending");
    DeviceGuard guard(h->cfg.device);
    if (term_frames_dev && !h->term_order)
        CK(cudaMalloc(&h->term_order, static_cast<size_t>(h->cfg.num_envs) * 8 * sizeof(int32_t)));
    h->term_frames = term_frames_dev;
    drop_graphs(h);
    return FFMP_OK;
}

int ffmp_launch_count(const ffmp_handle *h, uint64_t *out) {
    if (!h || !out) return fail(FFMP_ERR_ARG, "null argument");
    *out = h->launches;
    return FFMP_OK;
}

int ffmp_obs_slot(const ffmp_handle *h, int32_t *newest_slot) {
    if (!h || !newest_slot) return fail(FFMP_ERR_ARG, "null argument");
    *newest_slot = h->p;
    return FFMP_OK;
}

int ffmp_set_obs_slot(ffmp_handle *h, int32_t newest_slot) {
    if (!h) return fail(FFMP_ERR_ARG, "handle is null");
    if (newest_slot < 1 || newest_slot > (h->cfg.ring > 2 ? h->cfg.ring - 1 : 1)) return fail(FFMP_ERR_ARG, "newest_slot out of range");
    h->p = newest_slot;
    return FFMP_OK;
}

int ffmp_learner_input(ffmp_handle *h, void *out_dev, int32_t dtype, float scale, void *stream) {
    if (!h || !out_dev) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_learner_input");
    if (dtype != 0 && dtype != 1) return fail(FFMP_ERR_ARG, "dtype must be 0 (float32) or 1 (bfloat16)");
    if (reinterpret_cast<uintptr_t>(out_dev) % 16) return fail(FFMP_ERR_ARG, "out must be 16-byte aligned");
    DeviceGuard guard(h->cfg.device);
    ffmp::FeedArgs a{};
    a.N = h->cfg.num_envs; a.K = h->cfg.ring; a.W = h->cfg.window; a.slot_new = h->p; a.bf16 = dtype;
    a.scale = scale; a.frames = h->b.frames; a.out = out_dev;
    CK(ffmp::launch_learner_input(a, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return FFMP_OK;
}

int ffmp_scan(ffmp_handle *h, int32_t beams, float range_max, float *scan_dev, uint8_t *hit_dev, void *stream) {
    if (!h || !scan_dev) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_scan");
    if (beams <= 0 || beams > 65536 || !(range_max > 0.0f)) return fail(FFMP_ERR_ARG, "beams must be in [1,65536] and range_max > 0");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the scenario slot an env has just switched to may still be regenerating for a *later* episode only: the current
    // slot is complete by construction (run_tick waits for it), so no join is needed here
    ffmp::ScanArgs a{};
    a.n = h->cfg.num_envs; a.G = h->cfg.grid; a.beams = beams; a.range_max = range_max; a.flow_mode = 1;
    a.map = h->b.flow; a.state = h->b.state; a.S = h->cfg.slots; a.N = h->cfg.num_envs;
    a.scan = scan_dev; a.hit = hit_dev;
    CK(ffmp::launch_scan(a, st));
    h->launches += 1;
    return FFMP_OK;
}

int ffmp_op_scan(int32_t device, int32_t n, int32_t G, const uint8_t *map_dev, int32_t flow_mode, const float *pose_dev,
                 int32_t beams, float range_max, float *scan_dev, uint8_t *hit_dev, void *stream) {
    if (n < 0 || !map_dev || !pose_dev || !scan_dev) return fail(FFMP_ERR_ARG, "bad argument");
    if (G < 1 || G > 32768) return fail(FFMP_ERR_ARG, "G out of range");
    if (beams <= 0 || beams > 65536 || !(range_max > 0.0f)) return fail(FFMP_ERR_ARG, "beams must be in [1,65536] and range_max > 0");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    ffmp::ScanArgs a{};
    a.n = n; a.G = G; a.beams = beams; a.range_max = range_max; a.flow_mode = flow_mode ? 1 : 0;
    a.map = map_dev; a.pose = pose_dev; a.scan = scan_dev; a.hit = hit_dev;
    CK(ffmp::launch_scan(a, static_cast<cudaStream_t>(stream)));
    return FFMP_OK;
}

// ---- learner feed over peer memory (feed.cu) ------------------------------------------------------------------------
struct ffmp_feed {
    int device = 0, world = 1, rank = 0;
    size_t block = 0, slot_stride = 0, buffer_stride = 0;
    char *base = nullptr;                     // local allocation: header (4 KB) + 2 buffers x world slots
    char *peer[ffmp::FEED_MAX_WORLD];         // base of rank r's allocation as mapped here (own rank: base)
    bool opened[ffmp::FEED_MAX_WORLD];
    uint32_t seq = 0;                         // sequence number of the latest local push
    uint64_t launches = 0;

    static constexpr size_t HDR = 4096, FLAG0 = 0, ACK0 = 1024, TICKET = 2048, ERRW = 2112, WORD_STRIDE = 64;
    uint32_t *flag_in(int owner, int src) const { return reinterpret_cast<uint32_t *>(peer[owner] + FLAG0 + WORD_STRIDE * src); }
    uint32_t *ack_in(int owner, int consumer) const { return reinterpret_cast<uint32_t *>(peer[owner] + ACK0 + WORD_STRIDE * consumer); }
    uint32_t *ticket() const { return reinterpret_cast<uint32_t *>(base + TICKET); }
    uint32_t *error_word() const { return reinterpret_cast<uint32_t *>(base + ERRW); }
    char *slot_in(int owner, uint32_t sequence, int src) const {
        return peer[owner] + HDR + (sequence & 1u) * buffer_stride + static_cast<size_t>(src) * slot_stride;
    }
};

int ffmp_feed_create(int32_t device, int32_t world, int32_t rank, size_t block_bytes, ffmp_feed **out) {
    if (!out) return fail(FFMP_ERR_ARG, "out is null");
    *out = nullptr;
    if (world < 1 || world > ffmp::FEED_MAX_WORLD || rank < 0 || rank >= world || block_bytes == 0)
        return fail(FFMP_ERR_ARG, "world must be in [1,16], rank in [0,world), block_bytes > 0");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    ffmp_feed *f = new (std::nothrow) ffmp_feed();
    if (!f) return fail(FFMP_ERR_ARG, "out of host memory");
    f->device = device; f->world = world; f->rank = rank; f->block = block_bytes;
    f->slot_stride = align_up(block_bytes, 256);
    f->buffer_stride = f->slot_stride * static_cast<size_t>(world);
    for (int r = 0; r < ffmp::FEED_MAX_WORLD; ++r) { f->peer[r] = nullptr; f->opened[r] = false; }
    const size_t total = ffmp_feed::HDR + 2 * f->buffer_stride;
    void *p = nullptr;
    cudaError_t ce = cudaMalloc(&p, total);
    if (ce == cudaSuccess) ce = cudaMemset(p, 0, ffmp_feed::HDR);
    if (ce != cudaSuccess) {
        if (p) cudaFree(p);
        delete f;
        return fail(FFMP_ERR_CUDA, "feed allocation", ce);
    }
    f->base = static_cast<char *>(p);
    f->peer[rank] = f->base;
    *out = f;
    return FFMP_OK;
}

int ffmp_feed_handle(ffmp_feed *f, uint8_t *handle_out) {
    if (!f || !handle_out) return fail(FFMP_ERR_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "FFMP_IPC_HANDLE_BYTES");
    DeviceGuard guard(f->device);
    cudaIpcMemHandle_t hd;
    CK(cudaIpcGetMemHandle(&hd, f->base));
    std::memcpy(handle_out, &hd, sizeof(hd));
    return FFMP_OK;
}

int ffmp_feed_connect(ffmp_feed *f, int32_t peer_rank, const uint8_t *handle) {
    if (!f || !handle) return fail(FFMP_ERR_ARG, "null argument");
    if (peer_rank < 0 || peer_rank >= f->world) return fail(FFMP_ERR_ARG, "peer rank out of range");
    if (peer_rank == f->rank || f->opened[peer_rank]) return FFMP_OK;
    DeviceGuard guard(f->device);
    cudaIpcMemHandle_t hd;
    std::memcpy(&hd, handle, sizeof(hd));
    void *p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, hd, cudaIpcMemLazyEnablePeerAccess));
    f->peer[peer_rank] = static_cast<char *>(p);
    f->opened[peer_rank] = true;
    return FFMP_OK;
}

int ffmp_feed_info(const ffmp_feed *f, void **base_dev, size_t *slot_stride, size_t *buffer_stride, uint32_t *seq) {
    if (!f) return fail(FFMP_ERR_ARG, "feed is null");
    if (base_dev) *base_dev = f->base + ffmp_feed::HDR;
    if (slot_stride) *slot_stride = f->slot_stride;
    if (buffer_stride) *buffer_stride = f->buffer_stride;
    if (seq) *seq = f->seq;
    return FFMP_OK;
}

int ffmp_feed_push(ffmp_handle *h, ffmp_feed *f, uint32_t dest_mask, double timeout_s, void *stream) {
    if (!h || !f) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_feed_push");
    const ffmp_cfg &c = h->cfg;
    const size_t N = c.num_envs, W = c.window;
    if (f->block != N * (2 * W * W + 21)) return fail(FFMP_ERR_ARG, "feed block size does not match the env (N * (2 W^2 + 21))");
    if ((2 * W * W) % 16 || reinterpret_cast<uintptr_t>(h->b.frames) % 16) return fail(FFMP_ERR_ARG, "2 W^2 must be a multiple of 16");
    dest_mask &= f->world >= 32 ? 0xFFFFFFFFu : ((1u << f->world) - 1u);
    for (int r = 0; r < f->world; ++r)
        if (((dest_mask >> r) & 1u) && !f->peer[r]) return fail(FFMP_ERR_STATE, "destination rank is not connected (ffmp_feed_connect)");
    DeviceGuard guard(f->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const uint32_t seq = ++f->seq;
    // credit: the buffer of this parity was last used by push seq-2; every destination must have released it
    if (seq > 2 && dest_mask) {
        CK(ffmp::launch_feed_spin(reinterpret_cast<const uint32_t *>(f->base + ffmp_feed::ACK0), ffmp_feed::WORD_STRIDE / 4,
                                  dest_mask, seq - 2, timeout_s, f->error_word(), st));
        f->launches += 1;
    }
    ffmp::FeedPushArgs a{};
    a.N = c.num_envs; a.K = c.ring; a.W = c.window; a.slot_new = h->p;
    a.frames = h->b.frames; a.rel_goal = h->b.rel_goal; a.velocity = h->b.velocity; a.reward = h->b.reward; a.done = h->b.done;
    a.ticket = f->ticket(); a.seq = seq;
    int nd = 0;
    for (int r = 0; r < f->world; ++r)
        if ((dest_mask >> r) & 1u) {
            a.dst[nd] = reinterpret_cast<uint8_t *>(f->slot_in(r, seq, f->rank));
            a.flag[nd] = f->flag_in(r, f->rank);
            ++nd;
        }
    a.ndst = nd;
    CK(ffmp::launch_feed_push(a, st));
    f->launches += 1;
    return FFMP_OK;
}

int ffmp_pack_transitions(ffmp_handle *h, void *dst_dev, void *stream) {
    if (!h || !dst_dev) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->ready) return fail(FFMP_ERR_STATE, "ffmp_reset must be called before ffmp_pack_transitions");
    const ffmp_cfg &c = h->cfg;
    if ((2 * c.window * c.window) % 16 || reinterpret_cast<uintptr_t>(dst_dev) % 16) return fail(FFMP_ERR_ARG, "dst must be 16-byte aligned and 2 W^2 a multiple of 16");
    DeviceGuard guard(c.device);
    ffmp::FeedPushArgs a{};
    a.N = c.num_envs; a.K = c.ring; a.W = c.window; a.slot_new = h->p;
    a.frames = h->b.frames; a.rel_goal = h->b.rel_goal; a.velocity = h->b.velocity; a.reward = h->b.reward; a.done = h->b.done;
    a.ndst = 1; a.dst[0] = static_cast<uint8_t *>(dst_dev);
    CK(ffmp::launch_feed_push(a, static_cast<cudaStream_t>(stream)));
    h->launches += 1;
    return FFMP_OK;
}

int ffmp_feed_wait(ffmp_feed *f, uint32_t src_mask, double timeout_s, void **buffer_dev, void *stream) {
    if (!f) return fail(FFMP_ERR_ARG, "feed is null");
    if (f->seq == 0) return fail(FFMP_ERR_STATE, "ffmp_feed_push must be called before ffmp_feed_wait (SPMD: every rank pushes every step)");
    DeviceGuard guard(f->device);
    src_mask &= f->world >= 32 ? 0xFFFFFFFFu : ((1u << f->world) - 1u);
    if (src_mask) {
        CK(ffmp::launch_feed_spin(reinterpret_cast<const uint32_t *>(f->base + ffmp_feed::FLAG0), ffmp_feed::WORD_STRIDE / 4, src_mask,
                                  f->seq, timeout_s, f->error_word(), static_cast<cudaStream_t>(stream)));
        f->launches += 1;
    }
    if (buffer_dev) *buffer_dev = f->base + ffmp_feed::HDR + (f->seq & 1u) * f->buffer_stride;
    return FFMP_OK;
}

int ffmp_feed_release(ffmp_feed *f, uint32_t src_mask, void *stream) {
    if (!f) return fail(FFMP_ERR_ARG, "feed is null");
    DeviceGuard guard(f->device);
    src_mask &= f->world >= 32 ? 0xFFFFFFFFu : ((1u << f->world) - 1u);
    ffmp::FeedTargets t{};
    for (int r = 0; r < f->world; ++r) {
        t.word[r] = nullptr;
        if ((src_mask >> r) & 1u) {
            if (!f->peer[r]) return fail(FFMP_ERR_STATE, "source rank is not connected (ffmp_feed_connect)");
            t.word[r] = f->ack_in(r, f->rank);
        }
    }
    if (src_mask) {
        CK(ffmp::launch_feed_signal(t, src_mask, f->seq, static_cast<cudaStream_t>(stream)));
        f->launches += 1;
    }
    return FFMP_OK;
}

int ffmp_feed_error(ffmp_feed *f, uint32_t *out, void *stream) {
    if (!f || !out) return fail(FFMP_ERR_ARG, "null argument");
    DeviceGuard guard(f->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(out, f->error_word(), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return FFMP_OK;
}

int ffmp_feed_destroy(ffmp_feed *f) {
    if (!f) return FFMP_OK;
    DeviceGuard guard(f->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < f->world; ++r)
        if (f->opened[r] && f->peer[r]) cudaIpcCloseMemHandle(f->peer[r]);
    if (f->base) cudaFree(f->base);
    delete f;
    return FFMP_OK;
}

int ffmp_join(ffmp_handle *h, void *stream) {
    if (!h) return fail(FFMP_ERR_ARG, "handle is null");
    DeviceGuard guard(h->cfg.device);
    // episode ends of an unfinished group of ticks (regen_batch > 1) are handed to the regeneration now
    if (h->group_open) if (int rc = launch_regen(h, static_cast<cudaStream_t>(stream), nullptr, true)) return rc;
    for (int l = 0; l < h->nlist; ++l)
        if (h->regen_pending[l]) CK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), h->ev_regen[l], 0));
    return FFMP_OK;
}

int ffmp_error_word(ffmp_handle *h, uint32_t *out, void *stream) {
    if (!h || !out) return fail(FFMP_ERR_ARG, "null argument");
    if (!h->bound) return fail(FFMP_ERR_STATE, "not bound");
    DeviceGuard guard(h->cfg.device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CK(cudaMemcpyAsync(out, h->error_word(), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return FFMP_OK;
}

int ffmp_op_scenarios(int32_t device, int32_t n, int32_t G, uint32_t p_thresh, int32_t goal_mode, int32_t block_shift,
                      uint64_t seed, const uint32_t *env_gid_dev, const uint32_t *episode_dev, uint8_t *occ_dev,
                      uint32_t *scen_dev, void *stream) {
    if (n < 0 || !env_gid_dev || !episode_dev || !occ_dev || !scen_dev) return fail(FFMP_ERR_ARG, "bad argument");
    if (G < 16 || G > 1024 || (G % 4)) return fail(FFMP_ERR_ARG, "G must be a multiple of 4 in [16,1024]");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    ffmp::ScenarioArgs a{};
    a.env_gid = env_gid_dev; a.episode = episode_dev; a.count = n;
    a.G = G; a.goal_mode = goal_mode; a.block_shift = block_shift; a.slot_mode = 0; a.S = 1; a.N = n;
    a.p_thresh = p_thresh; a.seed = seed; a.occ = occ_dev; a.scen = scen_dev;
    CK(ffmp::launch_scenarios(a, n < 148 * 8 ? n : 148 * 8, static_cast<cudaStream_t>(stream)));
    return FFMP_OK;
}

size_t ffmp_op_flow_field_workspace(int32_t n, int32_t G) {
    if (n <= 0 || !ffmp::flow_field_supported(G)) return 0;
    const int maxg = ffmp::flow_field_max_grid(G);
    // 256-byte header (work counter, completion ticket) + the per-CTA plane scratch
    return 256 + align_up(static_cast<size_t>(n < maxg ? n : maxg) * ffmp::flow_field_scratch_words(G) * 4, 256);
}

int ffmp_op_flow_field(int32_t device, int32_t n, int32_t G, const uint8_t *occ_dev, const int32_t *goal_cells_dev,
                       int32_t *cost_dev, uint8_t *flow_dev, void *workspace_dev, void *stream) {
    if (n < 0 || !occ_dev || !goal_cells_dev || !flow_dev || !workspace_dev) return fail(FFMP_ERR_ARG, "bad argument");
    if (!ffmp::flow_field_supported(G)) return fail(FFMP_ERR_ARG, "G must be a multiple of 4 in [16,128] or a multiple of 32 in (128,512]");
    if (reinterpret_cast<uintptr_t>(occ_dev) % 16 || reinterpret_cast<uintptr_t>(flow_dev) % 16 ||
        reinterpret_cast<uintptr_t>(cost_dev) % 16)
        return fail(FFMP_ERR_ARG, "occ / cost / flow must be 16-byte aligned");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    ffmp::FlowArgs a{};
    a.count = n; a.G = G; a.slot_mode = 0; a.S = 1; a.N = n;
    a.occ = occ_dev; a.goal_cells = goal_cells_dev; a.cost = cost_dev; a.flow = flow_dev;
    if (reinterpret_cast<uintptr_t>(workspace_dev) % 16) return fail(FFMP_ERR_ARG, "workspace must be 16-byte aligned");
    CK(cudaMemsetAsync(workspace_dev, 0, 16, st));
    a.work = static_cast<uint32_t *>(workspace_dev);       // dynamic grid hand-out counter + completion ticket
    a.ticket = a.work + 1;
    a.hi_scratch = static_cast<uint32_t *>(workspace_dev) + 64;
    const int maxg = ffmp::flow_field_max_grid(G);
    const int grid = n < maxg ? n : maxg;
    CK(ffmp::launch_flow_field(a, grid, st));
    return FFMP_OK;
}

static int run_rewarder(int32_t device, ffmp::RewarderArgs a, void *stream) {
    if (a.n < 0 || !a.rel_goal || !a.is_first || !a.d_first || !a.reward || !a.done || !a.flags)
        return fail(FFMP_ERR_ARG, "bad argument");
    if (int rc = check_device(device)) return rc;
    DeviceGuard guard(device);
    CK(ffmp::launch_rewarder(a, static_cast<cudaStream_t>(stream)));
    return FFMP_OK;
}

int ffmp_op_rewarder(int32_t device, int32_t n, int32_t W, const int32_t *local_map_dev, const float *rel_goal_dev,
                     const uint8_t *is_first_dev, float *d_first_dev, float *reward_dev, uint8_t *done_dev,
                     uint8_t *flags_dev, void *stream) {
    if (W < 6 || !local_map_dev) return fail(FFMP_ERR_ARG, "bad argument");
    ffmp::RewarderArgs a{};
    a.n = n; a.W = W; a.local_map = local_map_dev; a.rel_goal = rel_goal_dev; a.is_first = is_first_dev;
    a.d_first = d_first_dev; a.reward = reward_dev; a.done = done_dev; a.flags = flags_dev;
    return run_rewarder(device, a, stream);
}

int ffmp_op_rewarder2(int32_t device, int32_t n, int32_t scan_len, const double *scan_dev, const float *rel_goal_dev,
                      const uint8_t *is_first_dev, float *d_first_dev, float *reward_dev, uint8_t *done_dev,
                      uint8_t *flags_dev, void *stream) {
    if (scan_len < 0 || !scan_dev) return fail(FFMP_ERR_ARG, "bad argument");
    ffmp::RewarderArgs a{};
    a.n = n; a.scan = scan_dev; a.scan_len = scan_len; a.rel_goal = rel_goal_dev; a.is_first = is_first_dev;
    a.d_first = d_first_dev; a.reward = reward_dev; a.done = done_dev; a.flags = flags_dev;
    return run_rewarder(device, a, stream);
}

int ffmp_op_reward_calculator(int32_t device, int32_t n, const float *rel_goal_dev, const uint8_t *given_flags_dev,
                              const uint8_t *is_first_dev, float *d_first_dev, float *reward_dev, uint8_t *done_dev,
                              uint8_t *flags_dev, void *stream) {
    if (!given_flags_dev) return fail(FFMP_ERR_ARG, "bad argument");
    ffmp::RewarderArgs a{};
    a.n = n; a.given_flags = given_flags_dev; a.rel_goal = rel_goal_dev; a.is_first = is_first_dev;
    a.d_first = d_first_dev; a.reward = reward_dev; a.done = done_dev; a.flags = flags_dev;
    return run_rewarder(device, a, stream);
}

}  // extern "C"
