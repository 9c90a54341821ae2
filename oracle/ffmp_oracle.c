/*
 * ffmp_oracle.c — CPU oracle for the gym_ffmp step()/reset() hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import, link or execute
 * this file.  Callers allowed: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline /
 * --impl reference legs.
 *
 * What it is: a deliberately dumb, sequential restatement of SPEC.md.
 *   - orc_ref_* functions restate the reference's own Python functions in fp64, line by line
 *     (citations on each); they are PINNED by tests/golden/ref_golden.json, which was produced
 *     by executing the unmodified reference (oracle/make_golden.py).
 *   - the remaining functions implement the [SPEC] parts the reference does not contain
 *     (hash map generator, BFS integration field, 8-neighbour argmin, fp32 unicycle, crop,
 *     step/reset sequencing, LiDAR scan synthesis).  For those rows parity is UNPINNED against the reference
 *     (SURVEY.md §0/§8c): they are pinned only by SPEC.md and by oracle<->CUDA equality.
 *
 * Build: gcc -O2 -ffp-contract=off -fno-fast-math -shared -fPIC (see oracle/build.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_INF 0x7FFFFFFF
#define RES 0.05f
#define INV_RES 20.0f
#define PI_F 3.14159274f
#define TWO_PI_F 6.28318548f
#define PI_2_F 1.57079637f
#define PI_4_F 0.785398185f

/* ------------------------------------------------------------------------------------------
 * Reference restatements (fp64, as CPython evaluates them)
 * ---------------------------------------------------------------------------------------- */

/* robot/config.py:25-58  RobotAction.cmd[i] */
void orc_ref_action(int id, double *v, double *w) {
    static const double V[4] = {0.0, 0.2, 0.4, 0.6};
    static const double Wt[7] = {-0.6, -0.4, -0.2, 0.0, 0.2, 0.4, 0.6};
    *v = V[id / 7];
    *w = Wt[id % 7];
}

/* ffmp.py:85-105  FFMP.is_collision(local_map_info); map is grid_num x grid_num ints */
int orc_ref_is_collision(const int32_t *local_map, int grid_num, double grid_size, double map_range,
                         double robot_rsize) {
    int collide = 0;
    for (int i = 0; i < grid_num && !collide; i++)
        for (int j = 0; j < grid_num; j++) {
            double xd = pow(i * grid_size - 0.5 * map_range, 2);
            double yd = pow(j * grid_size - 0.5 * map_range, 2);
            double dist = sqrt(xd + yd);
            if (dist <= robot_rsize && local_map[i * grid_num + j] > 0) { collide = 1; break; }
        }
    return collide;
}

/* ffmp.py:85-98  the footprint cell list itself; returns count, writes (i,j) pairs */
int orc_ref_footprint(int grid_num, double grid_size, double map_range, double robot_rsize, int32_t *ij) {
    int n = 0;
    for (int i = 0; i < grid_num; i++)
        for (int j = 0; j < grid_num; j++) {
            double xd = pow(i * grid_size - 0.5 * map_range, 2);
            double yd = pow(j * grid_size - 0.5 * map_range, 2);
            if (sqrt(xd + yd) <= robot_rsize) { ij[2 * n] = i; ij[2 * n + 1] = j; n++; }
        }
    return n;
}

/* ffmp.py:108-117  FFMP.is_collision2(scan); NaN encodes None, 0.0 is falsy */
int orc_ref_is_collision2(const double *scan, int n) {
    for (int i = 0; i < n; i++) {
        if (isnan(scan[i]) || scan[i] == 0.0) continue;
        if (scan[i] < 0.13) return 1;
    }
    return 0;
}

/* ffmp.py:120-127 */
int orc_ref_is_goal(double d) { return d < 0.5; }

/* ffmp.py:130-157; *pre is the module-global pre_relative_goal_dist */
double orc_ref_reward(double cur_dist, int is_collision, int is_goal, int is_first, double *pre) {
    double r_arr = 100.0 / 100.0, r_col = -100.0 / 100.0, r_s = -5.0 / 100.0, epsilon = 5.0 / 100.0;
    double r_g, r_c;
    if (is_first) *pre = cur_dist;
    if (is_goal) r_g = r_arr; else r_g = epsilon * (*pre - cur_dist);
    if (is_collision) r_c = r_col; else r_c = 0;
    return r_g + r_c + r_s;
}

/* ffmp.py:160-164 */
int orc_ref_is_done(int col, int goal) { return col || goal; }

/* train.py:167-172 */
double orc_ref_pi_to_pi(double a) {
    while (a >= M_PI) a = a - (2 * M_PI);
    while (a <= -M_PI) a = a + 2 * M_PI;
    return a;
}

/* train.py:174-180 */
void orc_ref_relative_goal(double gx, double gy, double x, double y, double yaw, double *out) {
    double rx = gx - x, ry = gy - y;
    out[0] = sqrt(rx * rx + ry * ry);
    out[1] = orc_ref_pi_to_pi(atan2(ry, rx) - yaw);
}

/* train.py:182-188 (pre pose handling is the caller's) */
void orc_ref_velocity(double x, double y, double yaw, double px, double py, double pyaw, double *out) {
    out[0] = sqrt(pow(x - px, 2) + pow(y - py, 2));
    out[1] = orc_ref_pi_to_pi(yaw - pyaw);
}

/* ------------------------------------------------------------------------------------------
 * [SPEC] fp32 elementary functions (SPEC.md §6).  volatile-free: relies on -ffp-contract=off
 * and SSE scalar float arithmetic (x86-64), one rounding per operation.
 * ---------------------------------------------------------------------------------------- */
float orc_pi_to_pi(float a) {
    while (a >= PI_F) a = a - TWO_PI_F;
    while (a <= -PI_F) a = a + TWO_PI_F;
    return a;
}

void orc_sincos(float a, float *s, float *c) {
    float q = rintf(a * 0.636619747f);
    float r = a - q * 1.5703125f;
    r = r - q * 4.837512969970703125e-4f;
    r = r - q * 7.54978995489188216e-8f;
    float z = r * r;
    float sp = ((((-1.9515295891e-4f * z) + 8.3321608736e-3f) * z - 1.6666654611e-1f) * z) * r + r;
    float cp = ((((2.443315711809948e-5f * z) - 1.388731625493765e-3f) * z + 4.166664568298827e-2f) * z) * z;
    cp = cp - 0.5f * z;
    cp = cp + 1.0f;
    int n = ((int)q) & 3;
    switch (n) {
    case 0: *s = sp; *c = cp; break;
    case 1: *s = cp; *c = -sp; break;
    case 2: *s = -sp; *c = -cp; break;
    default: *s = -cp; *c = sp; break;
    }
}

float orc_atan2(float y, float x) {
    if (x == 0.0f && y == 0.0f) return 0.0f;
    float ax = fabsf(x), ay = fabsf(y);
    int swap = ay > ax;
    float t = (swap ? ax : ay) / (swap ? ay : ax);
    float base, u;
    if (t > 0.414213568f) { base = PI_4_F; u = (t - 1.0f) / (t + 1.0f); }
    else { base = 0.0f; u = t; }
    float z = u * u;
    float p = ((((8.05374449538e-2f * z - 1.38776856032e-1f) * z + 1.99777106478e-1f) * z - 3.33329491539e-1f) * z) * u + u;
    float a = base + p;
    if (swap) a = PI_2_F - a;
    if (x < 0.0f) a = PI_F - a;
    if (y < 0.0f) a = -a;
    return a;
}

float orc_dist(float dx, float dy) { return sqrtf(dx * dx + dy * dy); }

/* ------------------------------------------------------------------------------------------
 * [SPEC] hash RNG + scenario generation (SPEC.md §3)
 * ---------------------------------------------------------------------------------------- */
uint32_t orc_mix32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}
uint32_t orc_key(uint64_t seed, uint32_t env_gid, uint32_t episode) {
    uint32_t k = orc_mix32((uint32_t)seed ^ 0x9E3779B9U);
    k = orc_mix32(k ^ (uint32_t)(seed >> 32));
    k = orc_mix32(k ^ env_gid);
    k = orc_mix32(k ^ episode);
    return k;
}
uint32_t orc_draw(uint32_t key, uint32_t stream, uint32_t t) {
    return orc_mix32(orc_mix32(key ^ stream) + t * 0x9E3779B1U);
}
#define S_START 0x53544152U
#define S_GOAL 0x474F414CU
#define S_YAW 0x59415721U

static void cell_of(uint32_t u, int G, int *i, int *j) {
    uint32_t span = (uint32_t)(G - 6);
    *i = 3 + (int)((u & 0xFFFFU) % span);
    *j = 3 + (int)((u >> 16) % span);
}

/* start[3] = x,y,yaw ; goal[2] = gx,gy ; cells[4] = si,sj,gi,gj */
void orc_scenario(uint64_t seed, uint32_t env_gid, uint32_t episode, int G, uint32_t p_thresh, int goal_mode,
                  int block_shift, uint8_t *occ, float *start, float *goal, int32_t *cells) {
    uint32_t key = orc_key(seed, env_gid, episode);
    int si, sj, gi, gj;
    const int D2 = 20 * 20;
    if (goal_mode == 0) {
        cell_of(orc_draw(key, S_START, 0), G, &si, &sj);
        for (uint32_t t = 0; t < 64; t++) {
            cell_of(orc_draw(key, S_GOAL, t), G, &gi, &gj);
            if ((gi - si) * (gi - si) + (gj - sj) * (gj - sj) >= D2) break;
        }
    } else {
        gi = G - 8; gj = G - 8;
        for (uint32_t t = 0; t < 64; t++) {
            cell_of(orc_draw(key, S_START, t), G, &si, &sj);
            if ((gi - si) * (gi - si) + (gj - sj) * (gj - sj) >= D2) break;
        }
    }
    float yaw = (float)(orc_draw(key, S_YAW, 0) >> 8) * (TWO_PI_F * 5.9604644775390625e-08f) - PI_F;
    yaw = orc_pi_to_pi(yaw);
    for (int i = 0; i < G; i++)
        for (int j = 0; j < G; j++) {
            uint8_t o;
            if (i == 0 || j == 0 || i == G - 1 || j == G - 1) o = 1;
            else if ((abs(i - si) <= 2 && abs(j - sj) <= 2) || (abs(i - gi) <= 2 && abs(j - gj) <= 2)) o = 0;
            else {
                uint32_t blk = ((uint32_t)(i >> block_shift) << 16) | (uint32_t)(j >> block_shift);
                o = orc_mix32(key + blk * 0x9E3779B1U) < p_thresh ? 1 : 0;
            }
            occ[i * G + j] = o;
        }
    start[0] = (float)si * RES; start[1] = (float)sj * RES; start[2] = yaw;
    goal[0] = (float)gi * RES; goal[1] = (float)gj * RES;
    cells[0] = si; cells[1] = sj; cells[2] = gi; cells[3] = gj;
}

/* ------------------------------------------------------------------------------------------
 * [SPEC] F1 integration field: queue BFS (SPEC.md §4)
 * ---------------------------------------------------------------------------------------- */
void orc_integration_field(const uint8_t *occ, int G, int gi, int gj, int32_t *cost) {
    int n = G * G;
    for (int c = 0; c < n; c++) cost[c] = ORC_INF;
    if (gi < 0 || gj < 0 || gi >= G || gj >= G || occ[gi * G + gj]) return;
    int32_t *queue = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    int head = 0, tail = 0;
    cost[gi * G + gj] = 0;
    queue[tail++] = gi * G + gj;
    static const int DI[4] = {1, -1, 0, 0}, DJ[4] = {0, 0, 1, -1};
    while (head < tail) {
        int c = queue[head++];
        int i = c / G, j = c % G;
        for (int k = 0; k < 4; k++) {
            int ni = i + DI[k], nj = j + DJ[k];
            if (ni < 0 || nj < 0 || ni >= G || nj >= G) continue;
            int nc = ni * G + nj;
            if (occ[nc] || cost[nc] != ORC_INF) continue;
            cost[nc] = cost[c] + 1;
            queue[tail++] = nc;
        }
    }
    free(queue);
}

/* [SPEC] F2 flow direction: explicit 8-neighbour scan (SPEC.md §4) */
void orc_flow_dir(const uint8_t *occ, const int32_t *cost, int G, uint8_t *dir) {
    static const int DI[8] = {1, 1, 0, -1, -1, -1, 0, 1};
    static const int DJ[8] = {0, 1, 1, 1, 0, -1, -1, -1};
    for (int i = 0; i < G; i++)
        for (int j = 0; j < G; j++) {
            int c = i * G + j;
            uint8_t d = 8;
            if (cost[c] != ORC_INF) {
                int32_t best = cost[c];
                for (int k = 0; k < 8; k++) {
                    int ni = i + DI[k], nj = j + DJ[k];
                    if (ni < 0 || nj < 0 || ni >= G || nj >= G) continue;
                    if (occ[ni * G + nj]) continue;
                    if (k & 1) {
                        /* side cells (i+di, j) and (i, j+dj): in-grid because n is in-grid */
                        if (occ[ni * G + j] || occ[i * G + nj]) continue;
                    }
                    if (cost[ni * G + nj] < best) { best = cost[ni * G + nj]; d = (uint8_t)k; }
                }
            }
            dir[c] = d;
        }
}

/* [SPEC] flow image (SPEC.md §5) */
void orc_flow_image(const uint8_t *occ, const uint8_t *dir, int G, uint8_t *flow) {
    for (int c = 0; c < G * G; c++) flow[c] = occ[c] ? 255 : (uint8_t)(dir[c] * 28);
}

/* occ -> cost, dir, flow in one call */
void orc_flow_field(const uint8_t *occ, int G, int gi, int gj, int32_t *cost, uint8_t *dir, uint8_t *flow) {
    orc_integration_field(occ, G, gi, gj, cost);
    orc_flow_dir(occ, cost, G, dir);
    orc_flow_image(occ, dir, G, flow);
}

/* [SPEC] crop (SPEC.md §5) */
void orc_crop(const uint8_t *flow, int G, int W, int ci, int cj, uint8_t *out) {
    for (int a = 0; a < W; a++)
        for (int b = 0; b < W; b++) {
            int i = ci - W / 2 + a, j = cj - W / 2 + b;
            out[a * W + b] = (i < 0 || j < 0 || i >= G || j >= G) ? 255 : flow[i * G + j];
        }
}

static int robot_cell(float x) { return (int)floorf(x * INV_RES + 0.5f); }

/* [SPEC] collision on the global flow image: 21 offsets di^2+dj^2 <= 6 */
int orc_collision(const uint8_t *flow, int G, int ci, int cj) {
    for (int di = -2; di <= 2; di++)
        for (int dj = -2; dj <= 2; dj++) {
            if (di * di + dj * dj > 6) continue;
            int i = ci + di, j = cj + dj;
            if (i < 0 || j < 0 || i >= G || j >= G) return 1;
            if (flow[i * G + j] == 255) return 1;
        }
    return 0;
}

/* [SPEC] §9 LiDAR scan synthesis: B beams through the grid (exact cell traversal), one fp32 rounding per operation.
 * map: flow image (flow_mode 1: 255 = occupied) or occupancy plane (flow_mode 0: non-zero = occupied).
 * Returns is_collision2 of the beam list (ffmp.py:108-117: a non-zero range below 0.13). */
static int scan_blocked(const uint8_t *map, int flow_mode, int G, int i, int j) {
    if (i < 0 || j < 0 || i >= G || j >= G) return 1;
    return flow_mode ? map[i * G + j] == 255 : map[i * G + j] != 0;
}

int orc_scan(const uint8_t *map, int flow_mode, int G, float x, float y, float yaw, int beams, float range_max,
             float *out) {
    const float inc = TWO_PI_F / (float)beams;
    const float u0 = x * INV_RES + 0.5f, v0 = y * INV_RES + 0.5f;
    const int i0 = (int)floorf(u0), j0 = (int)floorf(v0);
    int hit = 0;
    if (scan_blocked(map, flow_mode, G, i0, j0)) {
        for (int k = 0; k < beams; k++) out[k] = 0.0f;
        return 0;
    }
    const float fu = u0 - floorf(u0), fv = v0 - floorf(v0);
    const float max_t = range_max * INV_RES;
    for (int k = 0; k < beams; k++) {
        float s, c;
        orc_sincos(orc_pi_to_pi(yaw + (float)k * inc), &s, &c);
        int si, sj, i = i0, j = j0;
        float tdx, tmx, tdy, tmy;
        if (c > 0.0f) { si = 1; tdx = 1.0f / c; tmx = (1.0f - fu) * tdx; }
        else if (c < 0.0f) { si = -1; tdx = 1.0f / (-c); tmx = fu * tdx; }
        else { si = 0; tdx = tmx = INFINITY; }
        if (s > 0.0f) { sj = 1; tdy = 1.0f / s; tmy = (1.0f - fv) * tdy; }
        else if (s < 0.0f) { sj = -1; tdy = 1.0f / (-s); tmy = fv * tdy; }
        else { sj = 0; tdy = tmy = INFINITY; }
        float r;
        for (;;) {
            float t;
            if (tmx < tmy) { t = tmx; i += si; tmx = tmx + tdx; }
            else { t = tmy; j += sj; tmy = tmy + tdy; }
            if (t > max_t) { r = INFINITY; break; }
            if (scan_blocked(map, flow_mode, G, i, j)) { r = t * RES; break; }
        }
        out[k] = r;
        if (r != 0.0f && (double)r < 0.13) hit = 1;
    }
    return hit;
}

/* ------------------------------------------------------------------------------------------
 * [SPEC] batched env (SPEC.md §7); obs local_map is [N,2,W,W] = [older, newest]
 * ---------------------------------------------------------------------------------------- */
typedef struct {
    int N, G, W, max_steps, goal_mode, block_shift;
    uint32_t p_thresh, env_id_base;
    uint64_t seed;
    float dt;
    uint8_t *occ, *dir, *flow;   /* [N,G,G] current scenario */
    int32_t *cost;               /* [N,G,G] */
    float *pose, *goal;          /* [N,3], [N,2] */
    float *d_first, *ep_return;  /* [N] */
    int32_t *steps, *goal_cell;  /* [N], [N,2] */
    uint32_t *episode;           /* [N] */
    /* outputs */
    uint8_t *local_map;          /* [N,2,W,W] */
    uint8_t *term_local_map;     /* [N,2,W,W] terminal observation of the envs that finished in the last step (SPEC.md §7) */
    float *rel_goal, *velocity;  /* [N,2] */
    float *reward;               /* [N] */
    uint8_t *done, *flags;       /* [N] */
    float *term_rel_goal, *term_velocity, *fin_return; /* [N,2],[N,2],[N] */
    int32_t *fin_length;         /* [N] */
    uint32_t error_word;
} orc_env;

static void env_load_scenario(orc_env *e, int n) {
    size_t gg = (size_t)e->G * e->G;
    float start[3], goal[2];
    int32_t cells[4];
    orc_scenario(e->seed, e->env_id_base + (uint32_t)n, e->episode[n], e->G, e->p_thresh, e->goal_mode,
                 e->block_shift, e->occ + n * gg, start, goal, cells);
    orc_flow_field(e->occ + n * gg, e->G, cells[2], cells[3], e->cost + n * gg, e->dir + n * gg, e->flow + n * gg);
    memcpy(e->pose + 3 * n, start, sizeof start);
    memcpy(e->goal + 2 * n, goal, sizeof goal);
    e->goal_cell[2 * n] = cells[2]; e->goal_cell[2 * n + 1] = cells[3];
}

static void env_begin_episode(orc_env *e, int n) {
    size_t gg = (size_t)e->G * e->G, ww = (size_t)e->W * e->W;
    env_load_scenario(e, n);
    float x = e->pose[3 * n], y = e->pose[3 * n + 1], yaw = e->pose[3 * n + 2];
    float dx = e->goal[2 * n] - x, dy = e->goal[2 * n + 1] - y;
    float d = orc_dist(dx, dy);
    e->d_first[n] = d;
    e->steps[n] = 0;
    e->ep_return[n] = 0.0f;
    e->rel_goal[2 * n] = d;
    e->rel_goal[2 * n + 1] = orc_pi_to_pi(orc_atan2(dy, dx) - yaw);
    e->velocity[2 * n] = 0.0f; e->velocity[2 * n + 1] = 0.0f;
    uint8_t *lm = e->local_map + (size_t)n * 2 * ww;
    orc_crop(e->flow + n * gg, e->G, e->W, robot_cell(x), robot_cell(y), lm + ww);
    memcpy(lm, lm + ww, ww);
}

orc_env *orc_env_create(int N, int G, int W, int max_steps, int goal_mode, uint32_t p_thresh, uint64_t seed,
                        uint32_t env_id_base, float dt, int block_shift) {
    orc_env *e = (orc_env *)calloc(1, sizeof(orc_env));
    size_t gg = (size_t)G * G, ww = (size_t)W * W;
    e->N = N; e->G = G; e->W = W; e->max_steps = max_steps; e->goal_mode = goal_mode;
    e->p_thresh = p_thresh; e->seed = seed; e->env_id_base = env_id_base; e->dt = dt; e->block_shift = block_shift;
    e->occ = calloc(N * gg, 1); e->dir = calloc(N * gg, 1); e->flow = calloc(N * gg, 1);
    e->cost = calloc(N * gg, sizeof(int32_t));
    e->pose = calloc(N * 3, sizeof(float)); e->goal = calloc(N * 2, sizeof(float));
    e->d_first = calloc(N, sizeof(float)); e->ep_return = calloc(N, sizeof(float));
    e->steps = calloc(N, sizeof(int32_t)); e->goal_cell = calloc(N * 2, sizeof(int32_t));
    e->episode = calloc(N, sizeof(uint32_t));
    e->local_map = calloc(N * 2 * ww, 1);
    e->term_local_map = calloc(N * 2 * ww, 1);
    e->rel_goal = calloc(N * 2, sizeof(float)); e->velocity = calloc(N * 2, sizeof(float));
    e->reward = calloc(N, sizeof(float)); e->done = calloc(N, 1); e->flags = calloc(N, 1);
    e->term_rel_goal = calloc(N * 2, sizeof(float)); e->term_velocity = calloc(N * 2, sizeof(float));
    e->fin_return = calloc(N, sizeof(float)); e->fin_length = calloc(N, sizeof(int32_t));
    return e;
}

void orc_env_destroy(orc_env *e) {
    free(e->occ); free(e->dir); free(e->flow); free(e->cost); free(e->pose); free(e->goal);
    free(e->d_first); free(e->ep_return); free(e->steps); free(e->goal_cell); free(e->episode);
    free(e->local_map); free(e->term_local_map); free(e->rel_goal); free(e->velocity); free(e->reward); free(e->done); free(e->flags);
    free(e->term_rel_goal); free(e->term_velocity); free(e->fin_return); free(e->fin_length);
    free(e);
}

/* full reset: episode counters to 0 */
void orc_env_reset(orc_env *e) {
    for (int n = 0; n < e->N; n++) { e->episode[n] = 0; env_begin_episode(e, n); }
}

/* masked reset: masked envs abandon their episode and start the next one */
void orc_env_reset_masked(orc_env *e, const uint8_t *mask) {
    for (int n = 0; n < e->N; n++)
        if (mask[n]) { e->episode[n] += 1; env_begin_episode(e, n); }
}

static const float ACT_V[4] = {0.0f, 0.2f, 0.4f, 0.6f};
static const float ACT_W[7] = {-0.6f, -0.4f, -0.2f, 0.0f, 0.2f, 0.4f, 0.6f};

void orc_env_step(orc_env *e, const int64_t *actions) {
    size_t gg = (size_t)e->G * e->G, ww = (size_t)e->W * e->W;
    for (int n = 0; n < e->N; n++) {
        int64_t a = actions[n];
        if (a < 0 || a >= 28) { a = 3; e->error_word |= 1u; }
        float v = ACT_V[a / 7], w = ACT_W[a % 7];
        float x = e->pose[3 * n], y = e->pose[3 * n + 1], yaw = e->pose[3 * n + 2];
        float s, c;
        orc_sincos(yaw, &s, &c);
        float xn = x + (v * c) * e->dt;
        float yn = y + (v * s) * e->dt;
        float yawn = orc_pi_to_pi(yaw + w * e->dt);
        float dx = e->goal[2 * n] - xn, dy = e->goal[2 * n + 1] - yn;
        float d = orc_dist(dx, dy);
        float bearing = orc_pi_to_pi(orc_atan2(dy, dx) - yawn);
        float vl = orc_dist(xn - x, yn - y);
        float va = orc_pi_to_pi(yawn - yaw);
        int ci = robot_cell(xn), cj = robot_cell(yn);
        int col = orc_collision(e->flow + n * gg, e->G, ci, cj);
        int goal = d < 0.5f;
        float r = ((goal ? 1.0f : 0.05f * (e->d_first[n] - d)) + (col ? -1.0f : 0.0f)) + (-0.05f);
        e->steps[n] += 1;
        int trunc = e->steps[n] == e->max_steps;
        int done = col || goal || trunc;
        e->ep_return[n] += r;
        e->reward[n] = r;
        e->done[n] = (uint8_t)done;
        e->flags[n] = (uint8_t)((col ? 1 : 0) | (goal ? 2 : 0) | (trunc ? 4 : 0));
        e->term_rel_goal[2 * n] = d; e->term_rel_goal[2 * n + 1] = bearing;
        e->term_velocity[2 * n] = vl; e->term_velocity[2 * n + 1] = va;
        if (done) {
            /* terminal observation (a15, train.py:611-664: the last tick's obs): [previous newest frame, crop at the
             * terminal pose on the finished episode's flow image], taken before the next scenario replaces it */
            uint8_t *lm = e->local_map + (size_t)n * 2 * ww, *tm = e->term_local_map + (size_t)n * 2 * ww;
            memcpy(tm, lm + ww, ww);
            orc_crop(e->flow + n * gg, e->G, e->W, ci, cj, tm + ww);
            e->fin_return[n] = e->ep_return[n];
            e->fin_length[n] = e->steps[n];
            e->episode[n] += 1;
            env_begin_episode(e, n);
        } else {
            e->pose[3 * n] = xn; e->pose[3 * n + 1] = yn; e->pose[3 * n + 2] = yawn;
            e->rel_goal[2 * n] = d; e->rel_goal[2 * n + 1] = bearing;
            e->velocity[2 * n] = vl; e->velocity[2 * n + 1] = va;
            uint8_t *lm = e->local_map + (size_t)n * 2 * ww;
            memcpy(lm, lm + ww, ww);
            orc_crop(e->flow + n * gg, e->G, e->W, ci, cj, lm + ww);
        }
    }
}

/* field accessors for ctypes */
#define GETTER(T, name) T *orc_env_##name(orc_env *e) { return e->name; }
GETTER(uint8_t, occ) GETTER(uint8_t, dir) GETTER(uint8_t, flow) GETTER(int32_t, cost)
GETTER(float, pose) GETTER(float, goal) GETTER(float, d_first) GETTER(float, ep_return)
GETTER(int32_t, steps) GETTER(int32_t, goal_cell) GETTER(uint32_t, episode)
GETTER(uint8_t, local_map) GETTER(uint8_t, term_local_map) GETTER(float, rel_goal) GETTER(float, velocity) GETTER(float, reward)
GETTER(uint8_t, done) GETTER(uint8_t, flags) GETTER(float, term_rel_goal) GETTER(float, term_velocity)
GETTER(float, fin_return) GETTER(int32_t, fin_length)
uint32_t orc_env_error_word(orc_env *e) { return e->error_word; }
