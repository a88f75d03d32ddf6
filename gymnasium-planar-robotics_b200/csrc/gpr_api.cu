// gpr_api.cu — host side of the C ABI declared in include/gpr.h (the drop-in boundary).
//
// The handle owns the structure-of-arrays state and the small read-only tables derived from the config; all kernels are
// enqueued on the caller's stream.  Nothing here depends on torch.  There is no CPU fallback: every entry point either
// launches the CUDA kernels or fails with an error code.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <new>
#include <vector>

#include "gpr_launch.h"

using namespace gpr;

// ---------------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
#define CU(call)                                                                                        \
    do {                                                                                                \
        cudaError_t e_ = (call);                                                                        \
        if (e_ != cudaSuccess)                                                                          \
            return fail(e_ == cudaErrorMemoryAllocation ? GPR_ERR_OUT_OF_MEMORY : GPR_ERR_CUDA, "%s: %s", #call, \
                        cudaGetErrorString(e_));                                                        \
    } while (0)

// ---------------------------------------------------------------------------------------------------------------------
// handle
// ---------------------------------------------------------------------------------------------------------------------
struct gpr_handle {
    gpr_config cfg;
    int device = 0;
    int G = 1;  // lanes per environment
    bool noise = false;
    uint64_t seed = 0;
    uint64_t launches = 0;
    int obs_dim = 0, goal_dim = 0, action_dim = 0;
    // state
    double2 *pos = nullptr, *vel = nullptr, *acc = nullptr, *goal = nullptr;
    int32_t* elapsed = nullptr;
    uint32_t* rng = nullptr;
    uint8_t* needs_reset = nullptr;
    float* ep_return = nullptr;
    double* stats = nullptr;
    uint32_t* fail_count = nullptr;
    uint32_t* debug_errors = nullptr;  // [DBG_NUM_SLOTS], counted by GPR_DEBUG_BOUNDS builds
    unsigned long long* reset_list = nullptr;  // auto-reset work list (see planning_autoreset_kernel)
    unsigned long long* reset_count = nullptr;  // [2] control words (reported warps, reserved slots), then [2] uint32 cursors
    int parity = 0;
    int num_sms = 148;
    // pushing state
    double2* act = nullptr;        // [B] jerk integrator state
    double* mover_rot = nullptr;   // [B,3] cos yaw, sin yaw, yaw rate
    double* obj_pos = nullptr;     // [B,4] x, y, cos yaw, sin yaw
    double* obj_vel = nullptr;     // [B,3]
    float* push_warm = nullptr;    // [B,GPR_PUSH_WARM] warm-start state of the contact solve
    // tables
    double *cx = nullptr, *cy = nullptr, *c_wall = nullptr, *c_mover = nullptr;
    uint16_t* cell = nullptr;
    double quirk_rsum[2] = {0, 0};
    bool goal_dirty = true;  // desired_goal rows must all be written by the next step (GPR_OUT_GOAL_ON_CHANGE)
    const void* goal_ptr = nullptr;  // the desired_goal buffer the last step wrote into
    // per-kernel timing (gpr_kernel_times)
    bool timing = false;
    std::vector<cudaEvent_t> tev;  // triples: before step kernel, between, after auto-reset kernel
    size_t tev_used = 0;
    // staging for the *_host entry points
    cudaStream_t host_stream = nullptr;
    // ordering between the caller's stream(s) and the private host_stream (a non-blocking stream: it does not even order
    // with the legacy default stream): every call that enqueues work on a caller stream records `order_ev` there, and a
    // *_host call waits for it on host_stream before launching.  The other direction needs no event: *_host calls
    // synchronise host_stream before they return.
    cudaEvent_t order_ev = nullptr;
    bool order_pending = false;
    // compact transport of the sparse results of gpr_step_host (copy-engine route, see step_host_compact)
    bool compact_now = false;              // plan_args: hand the compact buffers to the kernels, do not write desired_goal rows
    const void* host_goal_synced = nullptr;  // the caller's HOST desired_goal buffer that holds every env's current goal
    bool in_host_call = false;             // gpr_step / gpr_reset were entered from a *_host call
    cudaEvent_t count_ev = nullptr;
    void* d_stage = nullptr;  // device: action + all outputs
    void* h_stage = nullptr;  // pinned mirror
    size_t stage_bytes = 0;
};

template <typename T>
static int dalloc(T** p, size_t n) {
    CU(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
    CU(cudaMemset(*p, 0, std::max<size_t>(n, 1) * sizeof(T)));
    return GPR_OK;
}

static uint16_t cell_code(const gpr_config& c, int i, int j) {
    auto L = [&](int a, int b) -> int {
        if (a < 0 || b < 0 || a >= c.num_tiles_x || b >= c.num_tiles_y) return 0;
        return c.layout[a * c.num_tiles_y + b] != 0;
    };
    auto in = [&](int a, int b) { return a >= 0 && b >= 0 && a < c.num_tiles_x && b < c.num_tiles_y; };
    uint32_t code = 0;
    if (L(i, j)) code |= CELL_T;
    if (L(i - 1, j)) code |= CELL_W;
    if (L(i + 1, j)) code |= CELL_E;
    if (L(i, j - 1)) code |= CELL_S;
    if (L(i, j + 1)) code |= CELL_N;
    if (L(i - 1, j - 1)) code |= CELL_SW;
    if (L(i - 1, j + 1)) code |= CELL_NW;
    if (L(i + 1, j - 1)) code |= CELL_SE;
    if (L(i + 1, j + 1)) code |= CELL_NE;
    bool full = i >= 1 && j >= 1 && i <= c.num_tiles_x - 2 && j <= c.num_tiles_y - 2;
    for (int a = -1; a <= 1 && full; ++a)
        for (int b = -1; b <= 1; ++b) full = full && L(i + a, j + b);
    if (full) code |= CELL_3X3;
    auto ic = [&](int di, int dj) { return L(i, j) && L(i + di, j) && L(i, j + dj) && in(i + di, j + dj) && !L(i + di, j + dj); };
    if (ic(+1, -1)) code |= CELL_IC_SE;
    if (ic(+1, +1)) code |= CELL_IC_NE;
    if (ic(-1, -1)) code |= CELL_IC_SW;
    if (ic(-1, +1)) code |= CELL_IC_NW;
    return (uint16_t)code;
}

static int next_pow2(int n) {
    int g = 1;
    while (g < n) g <<= 1;
    return g;
}

// ---------------------------------------------------------------------------------------------------------------------
// ABI: introspection
// ---------------------------------------------------------------------------------------------------------------------
extern "C" uint32_t gpr_config_bytes(void) { return (uint32_t)sizeof(gpr_config); }
extern "C" uint32_t gpr_abi_version(void) { return GPR_ABI_VERSION; }
extern "C" const char* gpr_last_error(void) { return g_err; }
extern "C" int gpr_obs_dim(const gpr_handle* h) { return h ? h->obs_dim : GPR_ERR_INVALID_ARG; }
extern "C" int gpr_goal_dim(const gpr_handle* h) { return h ? h->goal_dim : GPR_ERR_INVALID_ARG; }
extern "C" int gpr_action_dim(const gpr_handle* h) { return h ? h->action_dim : GPR_ERR_INVALID_ARG; }
extern "C" uint64_t gpr_launch_count(const gpr_handle* h) { return h ? h->launches : 0; }

// ---------------------------------------------------------------------------------------------------------------------
// create / destroy
// ---------------------------------------------------------------------------------------------------------------------
static int validate(const gpr_config* c) {
    if (!c) return fail(GPR_ERR_INVALID_ARG, "config is NULL");
    if (c->struct_bytes != sizeof(gpr_config) || c->abi_version != GPR_ABI_VERSION)
        return fail(GPR_ERR_ABI_MISMATCH, "gpr_config mismatch: caller %u bytes / ABI %u, library %zu bytes / ABI %d",
                    c->struct_bytes, c->abi_version, sizeof(gpr_config), GPR_ABI_VERSION);
    if (c->env_kind != GPR_ENV_PLANNING && c->env_kind != GPR_ENV_PUSHING)
        return fail(GPR_ERR_INVALID_ARG, "env_kind %d unknown", c->env_kind);
    if (c->num_envs <= 0) return fail(GPR_ERR_INVALID_ARG, "num_envs must be > 0");
    if (c->num_movers <= 0 || c->num_movers > GPR_MAX_MOVERS)
        return fail(GPR_ERR_INVALID_ARG, "num_movers must be in [1, %d]", GPR_MAX_MOVERS);
    if (c->env_kind == GPR_ENV_PUSHING && c->num_movers != 1)
        return fail(GPR_ERR_INVALID_ARG, "the pushing env has exactly one mover (pushing:196)");
    if (c->num_tiles_x <= 0 || c->num_tiles_y <= 0 || c->num_tiles_x > GPR_MAX_TILES_1D || c->num_tiles_y > GPR_MAX_TILES_1D)
        return fail(GPR_ERR_INVALID_ARG, "tile grid must be within [1, %d] per axis", GPR_MAX_TILES_1D);
    if (!(c->tile_half[0] > 0) || !(c->tile_half[1] > 0)) return fail(GPR_ERR_INVALID_ARG, "tile_half must be > 0");
    if (c->c_shape != GPR_SHAPE_CIRCLE && c->c_shape != GPR_SHAPE_BOX) return fail(GPR_ERR_INVALID_ARG, "c_shape unknown");
    if (c->num_cycles <= 0) return fail(GPR_ERR_INVALID_ARG, "num_cycles must be > 0");
    if (!(c->cycle_time > 0)) return fail(GPR_ERR_INVALID_ARG, "cycle_time must be > 0");
    if (!(c->v_max > 0) || !(c->a_max > 0) || !(c->j_max > 0)) return fail(GPR_ERR_INVALID_ARG, "v/a/j limits must be > 0");
    if (c->autoreset_mode < GPR_AUTORESET_OFF || c->autoreset_mode > GPR_AUTORESET_NEXT_STEP)
        return fail(GPR_ERR_INVALID_ARG, "autoreset_mode unknown");
    if (c->env_index_base < 0 || (uint64_t)c->env_index_base + (uint64_t)c->num_envs > (1ull << 32))
        return fail(GPR_ERR_INVALID_ARG, "global env indices must fit 32 bits");
    if (c->num_obstacles < 0 || c->num_obstacles > GPR_MAX_OBSTACLES) return fail(GPR_ERR_INVALID_ARG, "num_obstacles out of range");
    if (c->num_obstacles > 0 && c->env_kind != GPR_ENV_PLANNING)
        return fail(GPR_ERR_UNSUPPORTED, "static obstacles are a planning-env feature");
    for (int k = 0; k < c->num_obstacles; ++k) {
        if (!(c->obstacle_size[k][0] > 0) || (c->c_shape == GPR_SHAPE_BOX && !(c->obstacle_size[k][1] > 0)))
            return fail(GPR_ERR_INVALID_ARG, "obstacle sizes must be > 0");
        if (!std::isfinite(c->obstacle_vel[k][0]) || !std::isfinite(c->obstacle_vel[k][1]))
            return fail(GPR_ERR_INVALID_ARG, "obstacle velocities must be finite");
    }
    for (int m = 0; m < c->num_movers; ++m)
        for (int s = 0; s < 2; ++s) {
            if (!(c->c_wall[s][m][0] > 0) || !(c->c_mover[s][m][0] > 0))
                return fail(GPR_ERR_INVALID_ARG, "collision sizes must be > 0");
            // basic_envs.py:650 asserts when a shape is as wide as a tile; refuse such configs up front
            if (c->c_shape == GPR_SHAPE_CIRCLE && c->c_wall[s][m][0] >= std::min(c->tile_half[0], c->tile_half[1]))
                return fail(GPR_ERR_INVALID_ARG, "collision circle must be smaller than half a tile");
        }
    return GPR_OK;
}

extern "C" void gpr_destroy(gpr_handle* h) {
    if (!h) return;
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(h->device);
    void* ptrs[] = {h->pos, h->vel, h->acc, h->goal, h->elapsed, h->rng, h->needs_reset, h->ep_return, h->stats,
                    h->fail_count, h->debug_errors, h->reset_list, h->reset_count, h->act, h->mover_rot, h->obj_pos, h->obj_vel, h->push_warm, h->cx, h->cy, h->c_wall, h->c_mover,
                    h->cell, h->d_stage};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (cudaEvent_t ev : h->tev) cudaEventDestroy(ev);
    if (h->h_stage) cudaFreeHost(h->h_stage);
    if (h->order_ev) cudaEventDestroy(h->order_ev);
    if (h->count_ev) cudaEventDestroy(h->count_ev);
    if (h->host_stream) cudaStreamDestroy(h->host_stream);
    cudaSetDevice(prev);
    delete h;
}

extern "C" int gpr_create(const gpr_config* cfg, int device, gpr_handle** out_handle) {
    if (!out_handle) return fail(GPR_ERR_INVALID_ARG, "out_handle is NULL");
    *out_handle = nullptr;
    int rc = validate(cfg);
    if (rc != GPR_OK) return rc;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(GPR_ERR_NO_DEVICE, "no CUDA device visible: this library has no CPU fallback");
    if (device < 0 || device >= ndev) return fail(GPR_ERR_INVALID_ARG, "device %d out of range (have %d)", device, ndev);
    CU(cudaSetDevice(device));
    gpr_handle* h = new (std::nothrow) gpr_handle();
    if (!h) return fail(GPR_ERR_OUT_OF_MEMORY, "host allocation failed");
    h->cfg = *cfg;
    h->device = device;
    h->seed = cfg->seed;
    h->G = next_pow2(cfg->num_movers);
    h->noise = cfg->std_noise[0] != 0.0 || cfg->std_noise[1] != 0.0;  // (the object noise of push:565 has its own switch)
    const int N = cfg->num_movers, J = cfg->learn_jerk != 0;
    if (cfg->env_kind == GPR_ENV_PLANNING) {
        h->obs_dim = 2 * N * (1 + J);
        h->goal_dim = 2 * N;
        h->action_dim = 2 * N;
    } else {
        h->obs_dim = 2 * (2 + J);
        h->goal_dim = 2;
        h->action_dim = 2;
    }
    const size_t B = (size_t)cfg->num_envs, BN = B * (size_t)N;
#define TRY(x)                 \
    do {                       \
        rc = (x);              \
        if (rc != GPR_OK) {    \
            gpr_destroy(h);    \
            return rc;         \
        }                      \
    } while (0)
    TRY(dalloc(&h->pos, BN));
    TRY(dalloc(&h->vel, BN));
    TRY(dalloc(&h->acc, BN));
    TRY(dalloc(&h->goal, cfg->env_kind == GPR_ENV_PLANNING ? BN : B));
    TRY(dalloc(&h->elapsed, B));
    TRY(dalloc(&h->rng, B));
    TRY(dalloc(&h->needs_reset, B));
    TRY(dalloc(&h->ep_return, B));
    TRY(dalloc(&h->stats, 6));
    TRY(dalloc(&h->fail_count, 1));
    TRY(dalloc(&h->debug_errors, DBG_NUM_SLOTS));
    TRY(dalloc(&h->reset_list, B));
    CU(cudaMemset(h->reset_list, 0xFF, std::max<size_t>(B, 1) * sizeof(unsigned long long)));  // all ones: slot not published
    TRY(dalloc(&h->reset_count, 4));
    cudaDeviceGetAttribute(&h->num_sms, cudaDevAttrMultiProcessorCount, device);
    if (cfg->env_kind == GPR_ENV_PUSHING) {
        TRY(dalloc(&h->act, B));
        TRY(dalloc(&h->mover_rot, 3 * B));
        TRY(dalloc(&h->obj_pos, 4 * B));
        TRY(dalloc(&h->obj_vel, 3 * B));
        TRY(dalloc(&h->push_warm, GPR_PUSH_WARM * B));
    }
    TRY(dalloc(&h->cx, GPR_MAX_TILES_1D));
    TRY(dalloc(&h->cy, GPR_MAX_TILES_1D));
    TRY(dalloc(&h->c_wall, 2 * GPR_MAX_MOVERS * 2));
    TRY(dalloc(&h->c_mover, 2 * GPR_MAX_MOVERS * 2));
    TRY(dalloc(&h->cell, GPR_MAX_TILES_1D * GPR_MAX_TILES_1D));
#undef TRY
    std::vector<uint16_t> codes((size_t)cfg->num_tiles_x * cfg->num_tiles_y);
    for (int i = 0; i < cfg->num_tiles_x; ++i)
        for (int j = 0; j < cfg->num_tiles_y; ++j) codes[(size_t)i * cfg->num_tiles_y + j] = cell_code(*cfg, i, j);
    auto up = [&](void* dst, const void* src, size_t n) { return cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice); };
    cudaError_t e = up(h->cx, cfg->tile_cx, sizeof(cfg->tile_cx));
    if (e == cudaSuccess) e = up(h->cy, cfg->tile_cy, sizeof(cfg->tile_cy));
    if (e == cudaSuccess) e = up(h->c_wall, cfg->c_wall, sizeof(cfg->c_wall));
    if (e == cudaSuccess) e = up(h->c_mover, cfg->c_mover, sizeof(cfg->c_mover));
    if (e == cudaSuccess) e = up(h->cell, codes.data(), codes.size() * sizeof(uint16_t));
    if (e != cudaSuccess) {
        gpr_destroy(h);
        return fail(GPR_ERR_CUDA, "table upload: %s", cudaGetErrorString(e));
    }
    // basic:409 quirk: one threshold for all pairs = max over pairs of r_i + r_j
    for (int s = 0; s < 2; ++s) {
        double mx = -INFINITY;
        for (int i = 0; i < N - 1; ++i)
            for (int j = i + 1; j < N; ++j) mx = std::max(mx, cfg->c_mover[s][i][0] + cfg->c_mover[s][j][0]);
        h->quirk_rsum[s] = mx;
    }
    *out_handle = h;
    g_err[0] = 0;
    return GPR_OK;
}

// staging layout of the *_host entry points (device buffer and its pinned host mirror share it)
constexpr int kOutSlots = 13;  // pointers in gpr_outputs
constexpr int kCompact = 6;    // compact transport: final obs / ag / dg rows, new goal rows, env index, entry count
enum { C_FOBS = 0, C_FAG = 1, C_FDG = 2, C_GOAL = 3, C_INDEX = 4, C_COUNT = 5 };
struct StageLayout {
    size_t off_action, off[kOutSlots], bytes[kOutSlots], total;
    size_t coff[kCompact], crow[kCompact];  // offsets and bytes per row of the compact areas (B rows each; the count: one row)
};

static StageLayout stage_layout(const gpr_handle* h) {
    const size_t B = (size_t)h->cfg.num_envs;
    StageLayout L;
    size_t cur = 0;
    auto take = [&](size_t n) {
        size_t o = cur;
        cur += (n + 255) & ~(size_t)255;
        return o;
    };
    L.off_action = take(B * h->action_dim * sizeof(float));
    const size_t el = (h->cfg.output_flags & GPR_OUT_FLOAT64) ? sizeof(double) : sizeof(float);  // observation / goal arrays
    const size_t sz[kOutSlots] = {B * h->obs_dim * el, B * h->goal_dim * el, B * h->goal_dim * el,
                                  B * sizeof(float), B, B, B, B, B,
                                  B * h->obs_dim * el, B * h->goal_dim * el, B * h->goal_dim * el, B};
    for (int k = 0; k < kOutSlots; ++k) {
        L.bytes[k] = sz[k];
        L.off[k] = take(sz[k]);
    }
    const size_t crow[kCompact] = {h->obs_dim * el, h->goal_dim * el, h->goal_dim * el, h->goal_dim * el, sizeof(int32_t),
                                   sizeof(unsigned long long)};
    for (int k = 0; k < kCompact; ++k) {
        L.crow[k] = crow[k];
        L.coff[k] = take(k == C_COUNT ? crow[k] : B * crow[k]);
    }
    L.total = cur;
    return L;
}

// ---------------------------------------------------------------------------------------------------------------------
// argument packing and dispatch
// ---------------------------------------------------------------------------------------------------------------------
static LayoutArgs layout_args(const gpr_handle* h) {
    const gpr_config& c = h->cfg;
    LayoutArgs L;
    L.nx = c.num_tiles_x;
    L.ny = c.num_tiles_y;
    L.hx = c.tile_half[0];
    L.hy = c.tile_half[1];
    L.inv_wx = 1.0 / (2.0 * c.tile_half[0]);
    L.inv_wy = 1.0 / (2.0 * c.tile_half[1]);
    L.cx = h->cx;
    L.cy = h->cy;
    L.cell = h->cell;
    return L;
}

// GPR_NO_OVERLAP=1: launch the auto-reset kernel the ordinary way (after the step kernel has drained); for A/B timing
static bool no_overlap() {
    static const bool v = [] {
        const char* e = getenv("GPR_NO_OVERLAP");
        return e && e[0] == '1';
    }();
    return v;
}

static PlanArgs plan_args(const gpr_handle* h, const gpr_outputs* out) {
    const gpr_config& c = h->cfg;
    PlanArgs a;
    memset(&a, 0, sizeof(a));
    a.B = c.num_envs;
    a.N = c.num_movers;
    a.learn_jerk = c.learn_jerk != 0;
    a.num_cycles = c.num_cycles;
    a.max_episode_steps = c.max_episode_steps;
    a.autoreset = c.autoreset_mode;
    a.max_reset_attempts = c.max_reset_attempts;
    a.env_base = (uint32_t)c.env_index_base;
    a.seed = h->seed;
    a.dt = c.cycle_time;
    a.inv_dt = 1.0 / c.cycle_time;  // correctly rounded by the host: ddiv_rcp's premise
    a.v_max = c.v_max;
    a.a_max = c.a_max;
    a.j_max = c.j_max;
    a.act_lim = c.learn_jerk ? c.j_max : c.a_max;  // plan:257-259
    a.v_max2_lo = c.v_max * c.v_max * (1.0 - 1e-14);
    a.a_max2_lo = c.a_max * c.a_max * (1.0 - 1e-14);
    a.threshold = c.threshold_pos;
    a.min_goal_dist = c.min_goal_dist;
    for (int k = 0; k < 2; ++k) {
        a.min_xy[k] = c.min_xy_pos[k];
        a.span_xy[k] = c.max_xy_pos[k] - c.min_xy_pos[k];  // numpy uniform: low + (high-low)*u
    }
    // lazy-noise bounds: |normal| <= GPR_NORMAL_ABS_MAX, both components of a (x,y) noise vector share one Box-Muller
    // radius, so the vector norm is bounded by the same number.  6 > 5.77 and 12 > 2*5.77 leave slack for rounding.
    const double vb = 6.0 * c.std_noise[1] + 1e-12;
    a.v_lazy2 = (c.v_max - vb) > 0.0 ? (c.v_max - vb) * (c.v_max - vb) * (1.0 - 1e-14) : -1.0;
    a.pair_margin = h->noise ? 12.0 * c.std_noise[0] + 1e-12 : 0.0;
    a.wxf = (float)(2.0 * c.tile_half[0]);
    a.wyf = (float)(2.0 * c.tile_half[1]);
    // float32 prefilter slacks: float rounding of coordinates (<= nx*w) and of the thresholds, generously bounded
    const double extent = std::max(c.num_tiles_x * 2.0 * c.tile_half[0], c.num_tiles_y * 2.0 * c.tile_half[1]);
    const double fslack = 4e-6 * std::max(1.0, extent) + 2e-6;
    a.wall_delta = (float)((h->noise ? 6.0 * c.std_noise[0] * 1.01 : 0.0) + 1e-5 * std::max(2.0 * c.tile_half[0], 2.0 * c.tile_half[1]));
    a.pair_mgf[0] = (float)fslack;
    a.pair_mgf[1] = (float)(a.pair_margin + fslack);
    a.goal_lo2f = (float)std::pow(std::max(c.min_goal_dist - fslack, 0.0), 2);
    a.goal_hi2f = (float)std::pow(c.min_goal_dist + fslack, 2);
    a.minxf = (float)c.min_xy_pos[0];
    a.minyf = (float)c.min_xy_pos[1];
    a.spanxf = (float)(c.max_xy_pos[0] - c.min_xy_pos[0]);
    a.spanyf = (float)(c.max_xy_pos[1] - c.min_xy_pos[1]);
    // circle: one threshold for every pair when all radii are equal, or under the basic:409 broadcast quirk
    bool equal = true;
    for (int s = 0; s < 2; ++s)
        for (int m = 1; m < c.num_movers; ++m)
            equal = equal && c.c_mover[s][m][0] == c.c_mover[s][0][0] &&
                    (c.c_shape == GPR_SHAPE_CIRCLE || c.c_mover[s][m][1] == c.c_mover[s][0][1]);
    const bool quirk = c.reference_quirks != 0 && c.c_shape == GPR_SHAPE_CIRCLE;
    a.uniform_pairs = equal || quirk;  // box shape: all movers have one size
    for (int s = 0; s < 2; ++s) {
        const double t = (quirk && !equal) ? h->quirk_rsum[s] : c.c_mover[s][0][0] + c.c_mover[s][0][0];
        a.pair_t[s] = t;
        for (int nz = 0; nz < 2; ++nz) {
            const double mg = nz ? a.pair_margin : 0.0;
            a.band_lo2[s][nz] = (t - mg) > 0.0 ? (t - mg) * (t - mg) * (1.0 - 1e-14) : -1.0;
            a.band_hi2[s][nz] = (t + mg) * (t + mg) * (1.0 + 1e-14);
            const double mf = a.pair_mgf[nz];
            a.pair_lo2f[s][nz] = (t - mf) > 0.0 ? (float)((t - mf) * (t - mf)) : -1.f;
            a.pair_hi2f[s][nz] = (float)((t + mf) * (t + mf));
            a.pair_thif[s][nz] = (float)((t + mf) * 1.000001);
        }
    }
    // box shape: the quaternion noise (1 + n0*s, n1*s, n2*s, n3*s), |n| <= 5.77, rotates the rectangle by at most
    // |sin| <= 2*|qz|/|q|^2 < 14*s; the float32 wall screen widens its bounding rectangle by that much (s >= 5e-3 is
    // too large for the bound: a huge extent then makes the screen defer to the exact test every time)
    a.rot_extf = !h->noise ? 0.f : (c.std_noise[0] < 5e-3 ? (float)(14.0 * c.std_noise[0]) : 1e3f);
    a.inv_dtf = (float)((1.0 / c.cycle_time) * (1.0 - 1e-6));
    // static obstacles: float32 screen slack = position-noise bound + float rounding of coordinates up to the layout extent
    a.n_obst = c.num_obstacles;
    a.obst_delta = (float)((h->noise ? 6.0 * c.std_noise[0] * 1.01 : 0.0) + 4e-6 * std::max(1.0, extent) + 2e-6);
    double vmax = 0.0;
    for (int k = 0; k < c.num_obstacles; ++k) {
        a.obst[k][0] = c.obstacle_xy[k][0];
        a.obst[k][1] = c.obstacle_xy[k][1];
        a.obst[k][2] = c.obstacle_size[k][0];
        a.obst[k][3] = c.obstacle_size[k][1];
        a.obst_vel[k][0] = c.obstacle_vel[k][0];
        a.obst_vel[k][1] = c.obstacle_vel[k][1];
        vmax = std::max(vmax, std::hypot(c.obstacle_vel[k][0], c.obstacle_vel[k][1]));
    }
    a.obst_vmaxf = vmax > 0.0 ? (float)(vmax * 1.000001) + 1e-12f : 0.f;  // (rounded up: a bound)
    a.inv_wxf = (float)(1.0 / (2.0 * c.tile_half[0]));
    a.inv_wyf = (float)(1.0 / (2.0 * c.tile_half[1]));
    a.sigma_p = c.std_noise[0];
    a.sigma_v = c.std_noise[1];
    a.L = layout_args(h);
    a.c_wall = h->c_wall;
    a.c_mover = h->c_mover;
    a.pos = h->pos;
    a.vel = h->vel;
    a.acc = h->acc;
    a.goal = h->goal;
    a.elapsed = h->elapsed;
    a.rng = h->rng;
    a.needs_reset = h->needs_reset;
    a.ep_return = h->ep_return;
    a.stats = h->stats;
    a.fail_count = h->fail_count;
    a.debug_errors = h->debug_errors;
    a.reset_list = h->reset_list;
    a.reset_ctl = h->reset_count;
    a.reset_cursor = reinterpret_cast<uint32_t*>(h->reset_count + 2);
    a.overlap = (h->timing || no_overlap()) ? 0 : 1;  // (an event between the two launches would break the pairing)
    a.parity = h->parity;
    a.out_f64 = (c.output_flags & GPR_OUT_FLOAT64) != 0;
    if (out) a.out = *out;
    if (h->compact_now && h->d_stage) {
        const StageLayout L = stage_layout(h);
        char* ds = (char*)h->d_stage;
        a.compact_final_obs = (float*)(ds + L.coff[C_FOBS]);
        a.compact_final_ag = (float*)(ds + L.coff[C_FAG]);
        a.compact_final_dg = (float*)(ds + L.coff[C_FDG]);
        a.compact_goal = (float*)(ds + L.coff[C_GOAL]);
        a.compact_index = (int32_t*)(ds + L.coff[C_INDEX]);
    }
    return a;
}

static cudaError_t launch_plan(const gpr_handle* h, PlanKernel which, const PlanArgs& a, cudaStream_t s) {
    const bool box = h->cfg.c_shape == GPR_SHAPE_BOX;
    switch (h->G) {
        case 1: return launch_plan_g<1>(which, box, h->noise, a, h->num_sms, s);
        case 2: return launch_plan_g<2>(which, box, h->noise, a, h->num_sms, s);
        case 4: return launch_plan_g<4>(which, box, h->noise, a, h->num_sms, s);
        case 8: return launch_plan_g<8>(which, box, h->noise, a, h->num_sms, s);
        case 16: return launch_plan_g<16>(which, box, h->noise, a, h->num_sms, s);
        default: return launch_plan_g<32>(which, box, h->noise, a, h->num_sms, s);
    }
}

static PushArgs push_args(const gpr_handle* h, const gpr_outputs* out) {
    const gpr_config& c = h->cfg;
    PushArgs a;
    memset(&a, 0, sizeof(a));
    a.B = c.num_envs;
    a.learn_jerk = c.learn_jerk != 0;
    a.num_cycles = c.num_cycles;
    a.max_episode_steps = c.max_episode_steps;
    a.autoreset = c.autoreset_mode;
    a.max_reset_attempts = c.max_reset_attempts;
    a.env_base = (uint32_t)c.env_index_base;
    a.seed = h->seed;
    a.dt = c.cycle_time;
    a.inv_dt = 1.0 / c.cycle_time;  // correctly rounded by the host: ddiv_rcp's premise
    a.v_max = c.v_max;
    a.a_max = c.a_max;
    a.j_max = c.j_max;
    a.act_lim = c.learn_jerk ? c.j_max : c.a_max;  // push:243-245
    a.v_max2_lo = c.v_max * c.v_max * (1.0 - 1e-14);
    a.a_max2_lo = c.a_max * c.a_max * (1.0 - 1e-14);
    a.threshold = c.threshold_pos;
    for (int k = 0; k < 2; ++k) {
        a.min_xy[k] = c.min_xy_pos[k];
        a.span_xy[k] = c.max_xy_pos[k] - c.min_xy_pos[k];
        a.obj_min[k] = c.object_min_xy_pos[k];
        a.obj_span[k] = c.object_max_xy_pos[k] - c.object_min_xy_pos[k];
        for (int sft = 0; sft < 2; ++sft) a.c_wall[sft][k] = c.c_wall[sft][0][k];
    }
    a.min_mo_dist = c.min_mo_dist;
    {
        // lazy-noise / float-screen constants, the same construction as in plan_args
        const double vb = 6.0 * c.std_noise[1] + 1e-12;
        a.v_lazy2 = (c.v_max - vb) > 0.0 ? (c.v_max - vb) * (c.v_max - vb) * (1.0 - 1e-14) : -1.0;
        a.wxf = (float)(2.0 * c.tile_half[0]);
        a.wyf = (float)(2.0 * c.tile_half[1]);
        a.wall_delta = (float)((h->noise ? 6.0 * c.std_noise[0] * 1.01 : 0.0) + 1e-5 * std::max(2.0 * c.tile_half[0], 2.0 * c.tile_half[1]));
        a.inv_dtf = (float)((1.0 / c.cycle_time) * (1.0 - 1e-6));
    }
    a.sigma_p = c.std_noise[0];
    a.sigma_v = c.std_noise[1];
    a.sigma_obj = c.object_noise_xy;
    a.L = layout_args(h);
    gpr_push_params_from_config(&c, &a.P);  // planar physics parameters (include/gpr_push_physics.h)
    a.pos = h->pos;
    a.vel = h->vel;
    a.acc = h->acc;
    a.act = h->act;
    a.mover_rot = h->mover_rot;
    a.obj_pos = h->obj_pos;
    a.obj_vel = h->obj_vel;
    a.warm = h->push_warm;
    a.goal = h->goal;
    a.elapsed = h->elapsed;
    a.rng = h->rng;
    a.needs_reset = h->needs_reset;
    a.ep_return = h->ep_return;
    a.stats = h->stats;
    a.fail_count = h->fail_count;
    a.debug_errors = h->debug_errors;
    a.queue = h->reset_list;  // (the planning env's auto-reset work list: same size, unused by the pushing env)
    a.queue_ctl = h->reset_count;
    a.queue_cursor = reinterpret_cast<uint32_t*>(h->reset_count + 2);
    a.parity = h->parity;
    a.out_f64 = (c.output_flags & GPR_OUT_FLOAT64) != 0;
    if (out) a.out = *out;
    return a;
}

struct DeviceGuard {
    int prev = 0;
    explicit DeviceGuard(int dev) {
        cudaGetDevice(&prev);
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

// Work was enqueued on caller stream `s`: the next *_host call (private stream) must run after it.
static int note_caller_work(gpr_handle* h, cudaStream_t s) {
    // (no *_host call yet: nothing to order — the first one synchronises the device when it creates host_stream)
    if (!h->host_stream || s == h->host_stream) return GPR_OK;
    if (!h->order_ev) CU(cudaEventCreateWithFlags(&h->order_ev, cudaEventDisableTiming));
    CU(cudaEventRecord(h->order_ev, s));
    h->order_pending = true;
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// reset / step
// ---------------------------------------------------------------------------------------------------------------------
extern "C" int gpr_reset(gpr_handle* h, const uint8_t* reset_mask, int reseed, uint64_t seed, const double* inject_start,
                         const double* inject_goal, const double* inject_object, const gpr_outputs* out, void* stream) {
    if (!h) return fail(GPR_ERR_INVALID_ARG, "handle is NULL");
    DeviceGuard g(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    if (reseed) {
        // reference: reset(seed=...) re-creates np_random and rng_noise (basic_envs.py:1789-1791)
        h->seed = seed;
        CU(cudaMemsetAsync(h->rng, 0, sizeof(uint32_t) * (size_t)h->cfg.num_envs, s));
    }
    if (!out || out->desired_goal != h->goal_ptr) h->goal_dirty = true;  // new goals the steps' buffer does not see
    if (!h->in_host_call) h->host_goal_synced = nullptr;
    if (h->cfg.env_kind == GPR_ENV_PLANNING) {
        if (inject_object) return fail(GPR_ERR_INVALID_ARG, "inject_object is for the pushing env");
        PlanArgs a = plan_args(h, out);
        a.reset_mask = reset_mask;
        a.inject_start = reinterpret_cast<const double2*>(inject_start);
        a.inject_goal = reinterpret_cast<const double2*>(inject_goal);
        CU(launch_plan(h, PLAN_RESET, a, s));
    } else {
        PushArgs a = push_args(h, out);
        a.reset_mask = reset_mask;
        a.inject_start = reinterpret_cast<const double2*>(inject_start);
        a.inject_goal = reinterpret_cast<const double2*>(inject_goal);
        a.inject_object = reinterpret_cast<const double2*>(inject_object);
        CU(launch_push(PUSH_RESET, h->cfg.c_shape == GPR_SHAPE_BOX, h->noise, a, h->num_sms, s));
    }
    h->launches += 1;
    return note_caller_work(h, s);
}

extern "C" int gpr_step(gpr_handle* h, const float* action, const gpr_outputs* out, void* stream) {
    if (!h) return fail(GPR_ERR_INVALID_ARG, "handle is NULL");
    if (!action) return fail(GPR_ERR_INVALID_ARG, "action is NULL");
    if (!out) return fail(GPR_ERR_INVALID_ARG, "outputs struct is NULL");
    DeviceGuard g(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    cudaEvent_t* te = nullptr;
    if (h->timing) {
        while (h->tev.size() < h->tev_used + 3) {
            cudaEvent_t ev;
            CU(cudaEventCreate(&ev));
            h->tev.push_back(ev);
        }
        te = &h->tev[h->tev_used];
        h->tev_used += 3;
        CU(cudaEventRecord(te[0], s));
    }
    // GPR_OUT_GOAL_ON_CHANGE: rows of envs that are not reset are skipped — unless this buffer has not seen all of them yet
    int write_goal = !(h->cfg.output_flags & GPR_OUT_GOAL_ON_CHANGE) || h->goal_dirty || out->desired_goal != h->goal_ptr;
    if (h->compact_now) {
        write_goal = 0;  // (goal rows travel on the compact list; the caller's host buffer is known to be current)
        h->goal_dirty = true;  // ... but no device-side desired_goal buffer sees this step's new goals
    } else {
        h->goal_dirty = false;
        h->goal_ptr = out->desired_goal;
    }
    if (!h->in_host_call) h->host_goal_synced = nullptr;  // goals may change on the device without reaching a host buffer
    if (h->cfg.env_kind == GPR_ENV_PLANNING) {
        PlanArgs a = plan_args(h, out);
        a.action = reinterpret_cast<const float2*>(action);
        a.write_goal = write_goal;
        CU(launch_plan(h, PLAN_STEP, a, s));
        if (te) CU(cudaEventRecord(te[1], s));
        if (h->cfg.autoreset_mode != GPR_AUTORESET_OFF) {
            const cudaError_t ae = launch_plan(h, PLAN_AUTORESET, a, s);
            if (ae != cudaSuccess) {
                // the step kernel has published finished envs that nobody will consume: put the work list of this parity back
                // into its empty state (after the step kernel, in stream order) so the next step starts from a clean list;
                // the finished envs keep their terminal state and must be reset by the caller
                cudaMemsetAsync(h->reset_list, 0xFF, sizeof(unsigned long long) * (size_t)h->cfg.num_envs, s);
                cudaMemsetAsync(h->reset_count, 0, 4 * sizeof(unsigned long long), s);
                h->goal_dirty = true;
                return fail(GPR_ERR_CUDA, "auto-reset kernel launch: %s", cudaGetErrorString(ae));
            }
            h->parity ^= 1;
            h->launches += 1;
        }
    } else {
        PushArgs a = push_args(h, out);
        a.action = reinterpret_cast<const float2*>(action);
        a.write_goal = write_goal;
        CU(launch_push(PUSH_STEP, h->cfg.c_shape == GPR_SHAPE_BOX, h->noise, a, h->num_sms, s));
        if (te) CU(cudaEventRecord(te[1], s));
        CU(launch_push(PUSH_CONTACT, h->cfg.c_shape == GPR_SHAPE_BOX, h->noise, a, h->num_sms, s));
        h->parity ^= 1;
        h->launches += 1;
    }
    if (te) CU(cudaEventRecord(te[2], s));
    h->launches += 1;
    return note_caller_work(h, s);
}

extern "C" int gpr_kernel_times(gpr_handle* h, int enable, double* host_ms) {
    if (!h) return fail(GPR_ERR_INVALID_ARG, "handle is NULL");
    DeviceGuard g(h->device);
    if (host_ms) {
        double t_step = 0.0, t_reset = 0.0;
        for (size_t k = 0; k + 2 < h->tev_used + 0 && k + 3 <= h->tev_used; k += 3) {
            CU(cudaEventSynchronize(h->tev[k + 2]));
            float a = 0.f, b = 0.f;
            CU(cudaEventElapsedTime(&a, h->tev[k], h->tev[k + 1]));
            CU(cudaEventElapsedTime(&b, h->tev[k + 1], h->tev[k + 2]));
            t_step += a;
            t_reset += b;
        }
        host_ms[0] = t_step;
        host_ms[1] = t_reset;
        host_ms[2] = (double)(h->tev_used / 3);
    }
    h->tev_used = 0;
    h->timing = enable != 0;
    return GPR_OK;
}

// ---------------------------------------------------------------------------------------------------------------------
// host-buffer entry points (what a user of the reference calls: NumPy in, NumPy out)
// ---------------------------------------------------------------------------------------------------------------------
static int ensure_stage(gpr_handle* h) {
    if (h->d_stage) return GPR_OK;
    const StageLayout L = stage_layout(h);
    CU(cudaDeviceSynchronize());  // work enqueued on caller streams before the private stream existed
    CU(cudaStreamCreateWithFlags(&h->host_stream, cudaStreamNonBlocking));
    CU(cudaMalloc(&h->d_stage, L.total));
    CU(cudaMemset(h->d_stage, 0, L.total));
    CU(cudaMallocHost(&h->h_stage, L.total));
    h->stage_bytes = L.total;
    return GPR_OK;
}

static void** out_slot(gpr_outputs* o, int k) {
    void** slots[kOutSlots] = {(void**)&o->observation,      (void**)&o->achieved_goal,       (void**)&o->desired_goal,
                        (void**)&o->reward,           (void**)&o->terminated,          (void**)&o->truncated,
                        (void**)&o->is_success,       (void**)&o->mover_collision,     (void**)&o->wall_collision,
                        (void**)&o->final_observation, (void**)&o->final_achieved_goal, (void**)&o->final_desired_goal,
                        (void**)&o->other_collision};
    return slots[k];
}

// Page-locked host memory (torch pinned tensors, cudaHostAlloc / cudaHostRegister'ed arrays) is mapped into the device's
// address space under unified addressing: returns the device alias of `p`, or NULL for pageable memory.
static void* device_alias(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

// Where the kernels write each result of a *_host call:
//   page-locked destination -> the kernels store straight into it through its device alias (zero-copy: the stores of
//                              finished CTAs cross PCIe while other CTAs still compute; no copy is enqueued at all)
//   pageable destination    -> device staging, then one async copy into the handle's pinned mirror and a memcpy
struct HostRoute {
    gpr_outputs dev;   // pointers handed to the kernels
    bool staged[kOutSlots];   // field k goes through d_stage (copy engine afterwards)
    void* pinned[kOutSlots];  // staged field whose destination is page-locked: the copy engine writes it directly
};

// How the results of *_host calls reach page-locked caller buffers:
//   zero-copy  the kernels store straight into the buffers over PCIe (fastest on an otherwise idle host: B200, 65,536 envs,
//              192M vs 119M env-steps/s)
//   dma        device staging, then the copy engine writes the caller's buffers (faster when several GPUs write into one
//              host at once — 8 ranks: 497M vs 436M env-steps/s in total: SM-issued PCIe writes of 8 GPUs contend in the host)
// GPR_HOST_IO=zerocopy | dma forces one; the default `auto` takes dma when this process is one of several ranks on the host
// (LOCAL_WORLD_SIZE / WORLD_SIZE > 1, as torchrun exports them) and zero-copy otherwise.  Read once per process.
static bool host_io_dma() {
    static const bool v = [] {
        const char* e = getenv("GPR_HOST_IO");
        if (e && strcmp(e, "dma") == 0) return true;
        if (e && (strcmp(e, "zerocopy") == 0 || strcmp(e, "zc") == 0)) return false;
        const char* lw = getenv("LOCAL_WORLD_SIZE");
        const char* w = lw ? lw : getenv("WORLD_SIZE");
        return w && atoi(w) > 4;  // measured (profiles/r2_bench_multi_gpu.txt): zero-copy wins at 2 and 4 ranks, the compact copy-engine route at 8
    }();
    return v;
}

// GPR_HOST_COMPACT=0: the copy-engine route moves the dense final_* / desired_goal arrays instead of compact lists (A/B)
static bool host_io_no_compact() {
    static const bool v = [] {
        const char* e = getenv("GPR_HOST_COMPACT");
        return e && e[0] == '0';
    }();
    return v;
}

static HostRoute route_outputs(gpr_handle* h, const StageLayout& L, const gpr_outputs* host_out) {
    HostRoute r;
    gpr_outputs ho = *host_out;
    memset(&r, 0, sizeof(r));
    for (int k = 0; k < kOutSlots; ++k) {
        void* dst = *out_slot(&ho, k);
        if (!dst) continue;
        void* alias = device_alias(dst);
        r.pinned[k] = alias ? dst : nullptr;
        if (alias && host_io_dma()) alias = nullptr;
        r.staged[k] = alias == nullptr;
        *out_slot(&r.dev, k) = alias ? alias : (void*)((char*)h->d_stage + L.off[k]);
    }
    return r;
}

static int finish_host_call(gpr_handle* h, const StageLayout& L, const HostRoute& r, const gpr_outputs* host_out) {
    gpr_outputs ho = *host_out;
    char* hs = (char*)h->h_stage;
    char* ds = (char*)h->d_stage;
    for (int k = 0; k < kOutSlots; ++k)
        if (r.staged[k])
            CU(cudaMemcpyAsync(r.pinned[k] ? r.pinned[k] : (void*)(hs + L.off[k]), ds + L.off[k], L.bytes[k], cudaMemcpyDeviceToHost,
                               h->host_stream));
    CU(cudaStreamSynchronize(h->host_stream));  // results (zero-copy stores included) are visible to the host after this
    for (int k = 0; k < kOutSlots; ++k)
        if (r.staged[k] && !r.pinned[k]) memcpy(*out_slot(&ho, k), hs + L.off[k], L.bytes[k]);
    return GPR_OK;
}

// gpr_reset / gpr_step / gpr_set_state on a caller stream, then a *_host call: host_stream waits for the caller's work
static int order_host_stream(gpr_handle* h) {
    if (h->order_pending) {
        CU(cudaStreamWaitEvent(h->host_stream, h->order_ev, 0));
        h->order_pending = false;
    }
    return GPR_OK;
}

// COMPACT TRANSPORT (copy-engine route, planning env, SAME_STEP auto-reset).  Of the 201 bytes per env of result buffers only
// the dense rows (observation, achieved_goal, reward, flags: 74 B with 4 movers) change for every env; final_* and the new
// desired_goal concern the envs that finished — a third of them with random actions.  The kernels write those rows to
// row `slot` of compact device buffers (slot = the env's position on the auto-reset work list), the copy engine moves the
// dense arrays plus the first `count` rows of the compact ones, and this function scatters them into the caller's arrays:
// the caller sees exactly what the other routes deliver.  With 8 ranks on one host the copy engines together sustain
// 175-193 GB/s into host memory (tools/pcie_bw.py under torchrun) while SM-issued zero-copy stores of 8 GPUs get ~58 GB/s
// (profiles/r2_bench_8gpu.txt): this route moves ~104 B/env at the former rate.
static int step_host_compact(gpr_handle* h, const StageLayout& L, const float* dev_action, const gpr_outputs* host_out) {
    gpr_outputs ho = *host_out;
    char* hs = (char*)h->h_stage;
    char* ds = (char*)h->d_stage;
    if (!h->count_ev) CU(cudaEventCreateWithFlags(&h->count_ev, cudaEventDisableTiming));
    // dense slots through device staging; the sparse ones (final_*, desired_goal) are not written in place at all
    HostRoute r;
    memset(&r, 0, sizeof(r));
    const int sparse[4] = {2, 9, 10, 11};  // desired_goal, final_observation, final_achieved_goal, final_desired_goal
    for (int k = 0; k < kOutSlots; ++k) {
        void* dst = *out_slot(&ho, k);
        if (!dst || k == sparse[0] || k == sparse[1] || k == sparse[2] || k == sparse[3]) continue;
        r.pinned[k] = device_alias(dst) ? dst : nullptr;
        r.staged[k] = true;
        *out_slot(&r.dev, k) = (void*)(ds + L.off[k]);
    }
    const int par = h->parity;  // the work-list buffer this step uses (gpr_step flips it)
    h->compact_now = true;
    h->in_host_call = true;
    int rc = gpr_step(h, dev_action, &r.dev, h->host_stream);
    h->compact_now = false;
    h->in_host_call = false;
    if (rc != GPR_OK) return rc;
    // the entry count first (the list copies are sized by it), then the dense arrays behind it in the same stream
    unsigned long long* hcount = (unsigned long long*)(hs + L.coff[C_COUNT]);
    CU(cudaMemcpyAsync(hcount, h->reset_count + par, sizeof(unsigned long long), cudaMemcpyDeviceToHost, h->host_stream));
    CU(cudaEventRecord(h->count_ev, h->host_stream));
    for (int k = 0; k < kOutSlots; ++k)
        if (r.staged[k])
            CU(cudaMemcpyAsync(r.pinned[k] ? r.pinned[k] : (void*)(hs + L.off[k]), ds + L.off[k], L.bytes[k], cudaMemcpyDeviceToHost,
                               h->host_stream));
    CU(cudaEventSynchronize(h->count_ev));
    const size_t count = (size_t)(*hcount & 0xffffffffull);
    if (count > (size_t)h->cfg.num_envs) return fail(GPR_ERR_CUDA, "work-list count %zu exceeds num_envs", count);
    void* const sdst[4] = {ho.final_observation, ho.final_achieved_goal, ho.final_desired_goal, ho.desired_goal};
    // the caller takes the terminal observations as a list (gpr_outputs.final_index): the copy engine writes the compact rows
    // straight into its arrays (page-locked) and nothing but the new goals is scattered on the host
    const bool list_out = ho.final_index != nullptr && ho.final_count != nullptr;
    bool direct[4] = {false, false, false, false};
    if (count) {
        CU(cudaMemcpyAsync(hs + L.coff[C_INDEX], ds + L.coff[C_INDEX], count * L.crow[C_INDEX], cudaMemcpyDeviceToHost, h->host_stream));
        for (int c = 0; c < 4; ++c) {
            if (!sdst[c]) continue;
            direct[c] = list_out && c < 3 && device_alias(sdst[c]) != nullptr;
            CU(cudaMemcpyAsync(direct[c] ? sdst[c] : (void*)(hs + L.coff[c]), ds + L.coff[c], count * L.crow[c], cudaMemcpyDeviceToHost,
                               h->host_stream));
        }
    }
    CU(cudaStreamSynchronize(h->host_stream));
    for (int k = 0; k < kOutSlots; ++k)
        if (r.staged[k] && !r.pinned[k]) memcpy(*out_slot(&ho, k), hs + L.off[k], L.bytes[k]);
    const int32_t* idx = (const int32_t*)(hs + L.coff[C_INDEX]);
    for (size_t s_ = 0; s_ < count; ++s_)  // (the scatter below indexes the caller's arrays with these)
        if ((uint32_t)idx[s_] >= (uint32_t)h->cfg.num_envs) return fail(GPR_ERR_CUDA, "work-list entry %zu names env %d", s_, (int)idx[s_]);
    if (list_out) {
        memcpy(ho.final_index, idx, count * sizeof(int32_t));
        *ho.final_count = (uint32_t)count;
    }
    // scatter: row s of the compact arrays belongs to env index[s]
    for (int c = 0; c < 4; ++c) {
        if (!sdst[c]) continue;
        if (list_out && c < 3) {  // list form: rows stay compact
            if (!direct[c]) memcpy(sdst[c], hs + L.coff[c], count * L.crow[c]);
            continue;
        }
        const size_t rb = L.crow[c];
        const char* src = hs + L.coff[c];
        char* dst = (char*)sdst[c];
        if (rb % 8 == 0) {  // rows are whole (x, y) pairs: fixed-size word moves instead of a memcpy call per row
            const size_t w = rb / 8;
            const uint64_t* s8 = (const uint64_t*)src;
            for (size_t s_ = 0; s_ < count; ++s_) {
                uint64_t* d8 = (uint64_t*)(dst + (size_t)idx[s_] * rb);
                for (size_t q = 0; q < w; ++q) d8[q] = s8[s_ * w + q];
            }
        } else {
            for (size_t s_ = 0; s_ < count; ++s_) memcpy(dst + (size_t)idx[s_] * rb, src + s_ * rb, rb);
        }
    }
    return GPR_OK;
}

extern "C" int gpr_step_host(gpr_handle* h, const float* host_action, const gpr_outputs* host_out) {
    if (!h || !host_action || !host_out) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    int rc = ensure_stage(h);
    if (rc != GPR_OK) return rc;
    const StageLayout L = stage_layout(h);
    const size_t abytes = (size_t)h->cfg.num_envs * h->action_dim * sizeof(float);
    // action: a page-locked caller buffer is read by the kernel in place (each lane loads its float2 once, coalesced);
    // a pageable one is first copied into the pinned mirror, which the kernel then reads the same way
    const float* dev_action = (const float*)device_alias(host_action);
    if (!dev_action) {
        memcpy((char*)h->h_stage + L.off_action, host_action, abytes);
        dev_action = (const float*)device_alias((char*)h->h_stage + L.off_action);
        if (!dev_action) return fail(GPR_ERR_CUDA, "pinned staging buffer has no device alias");
    }
    // Measured on B200 (planning4, 65,536 envs): a copy-engine transfer of the action ahead of the kernel is slower end to
    // end than letting the kernel read the pinned buffer in place (158M vs 166M env-steps/s); splitting the step into 2-8
    // env chunks on separate streams (with or without per-chunk copy-engine transfers) changes nothing (+-3%).  What the
    // host-I/O step pays for is the result traffic (tools/sm_store_bw.cu): SM stores to pinned memory stream at 50 GB/s as
    // whole 32-byte sectors (the copy engine: 55), every write transaction carries ~24 B of overhead, and each FRAGMENT of
    // a partially written sector is a transaction of its own (0.8 G/s) — hence the CTA-coalesced flag / reward stores of
    // the step kernel, the one-store rows and the filler rows of the pushing kernels.
    rc = order_host_stream(h);
    if (rc != GPR_OK) return rc;
    // compact transport: the copy-engine route of a planning env with SAME_STEP auto-reset whose caller keeps handing in the
    // desired_goal buffer that already holds every env's current goal (the previous host call made it so)
    if (host_io_dma() && !host_io_no_compact() && h->cfg.env_kind == GPR_ENV_PLANNING && h->cfg.autoreset_mode == GPR_AUTORESET_SAME_STEP &&
        (h->cfg.output_flags & GPR_OUT_GOAL_ON_CHANGE) && host_out->desired_goal && h->host_goal_synced == host_out->desired_goal)
        return step_host_compact(h, L, dev_action, host_out);
    const HostRoute r = route_outputs(h, L, host_out);
    h->in_host_call = true;
    rc = gpr_step(h, dev_action, &r.dev, h->host_stream);
    h->in_host_call = false;
    if (rc != GPR_OK) return rc;
    rc = finish_host_call(h, L, r, host_out);
    // every env's goal is in the caller's buffer now if this call wrote all rows or the buffer was current already
    if (rc == GPR_OK && host_out->desired_goal) h->host_goal_synced = host_out->desired_goal;
    if (host_out->final_count) *host_out->final_count = 0xffffffffu;  // final_* rows were written densely (row = env)
    return rc;
}

extern "C" int gpr_reset_host(gpr_handle* h, int reseed, uint64_t seed, const gpr_outputs* host_out) {
    if (!h || !host_out) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    int rc = ensure_stage(h);
    if (rc != GPR_OK) return rc;
    const StageLayout L = stage_layout(h);
    const HostRoute r = route_outputs(h, L, host_out);
    rc = order_host_stream(h);
    if (rc != GPR_OK) return rc;
    h->in_host_call = true;
    rc = gpr_reset(h, nullptr, reseed, seed, nullptr, nullptr, nullptr, &r.dev, h->host_stream);
    h->in_host_call = false;
    if (rc != GPR_OK) return rc;
    rc = finish_host_call(h, L, r, host_out);
    if (rc == GPR_OK) h->host_goal_synced = host_out->desired_goal;  // a full reset wrote every env's goal
    return rc;
}

// ---------------------------------------------------------------------------------------------------------------------
// state access, HER rewards, statistics
// ---------------------------------------------------------------------------------------------------------------------
static int copy_state(gpr_handle* h, const gpr_state* st, bool to_handle, cudaStream_t s) {
    const size_t B = (size_t)h->cfg.num_envs, BN = B * (size_t)h->cfg.num_movers;
    const bool push = h->cfg.env_kind == GPR_ENV_PUSHING;
    struct Item {
        void* mine;
        void* theirs;
        size_t bytes;
    } items[] = {
        {h->pos, st->pos, BN * sizeof(double2)},
        {h->vel, st->vel, BN * sizeof(double2)},
        {h->acc, st->acc, BN * sizeof(double2)},
        {h->goal, st->goal, (push ? B : BN) * sizeof(double2)},
        {h->elapsed, st->elapsed_steps, B * sizeof(int32_t)},
        {h->rng, st->rng_counter, B * sizeof(uint32_t)},
        {h->act, st->act, B * sizeof(double2)},
        {h->mover_rot, st->mover_rot, 3 * B * sizeof(double)},
        {h->obj_pos, st->object_pos, 4 * B * sizeof(double)},
        {h->obj_vel, st->object_vel, 3 * B * sizeof(double)},
        {h->needs_reset, st->needs_reset, B * sizeof(uint8_t)},
        {h->ep_return, st->episode_return, B * sizeof(float)},
        {h->push_warm, st->contact_warm, GPR_PUSH_WARM * B * sizeof(float)},
    };
    for (const Item& it : items) {
        if (!it.theirs) continue;
        if (!it.mine) return fail(GPR_ERR_INVALID_ARG, "state field not present for this env kind");
        CU(cudaMemcpyAsync(to_handle ? it.mine : it.theirs, to_handle ? it.theirs : it.mine, it.bytes, cudaMemcpyDeviceToDevice, s));
    }
    return GPR_OK;
}

extern "C" int gpr_get_state(gpr_handle* h, const gpr_state* dst, void* stream) {
    if (!h || !dst) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    return copy_state(h, dst, false, (cudaStream_t)stream);
}

extern "C" int gpr_set_state(gpr_handle* h, const gpr_state* src, void* stream) {
    if (!h || !src) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    if (src->goal) h->goal_dirty = true;
    if (src->goal) h->host_goal_synced = nullptr;
    int rc = copy_state(h, src, true, (cudaStream_t)stream);
    return rc != GPR_OK ? rc : note_caller_work(h, (cudaStream_t)stream);
}

extern "C" int gpr_get_seed(const gpr_handle* h, uint64_t* seed) {
    if (!h || !seed) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    *seed = h->seed;
    return GPR_OK;
}

extern "C" int gpr_set_seed(gpr_handle* h, uint64_t seed) {
    if (!h) return fail(GPR_ERR_INVALID_ARG, "handle is NULL");
    h->seed = seed;
    return GPR_OK;
}

extern "C" int gpr_compute_reward(gpr_handle* h, int batch, const float* achieved, const float* desired,
                                  const uint8_t* mover_collision, const uint8_t* wall_collision, float* reward,
                                  uint8_t* terminated, void* stream) {
    if (!h || !achieved || !desired) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    if (batch < 0) return fail(GPR_ERR_INVALID_ARG, "batch < 0");
    if (batch == 0) return GPR_OK;
    DeviceGuard g(h->device);
    CU(launch_compute_reward(h->cfg.env_kind, h->cfg.num_movers, batch, h->cfg.threshold_pos, achieved, desired, false, mover_collision,
                             wall_collision, reward, terminated, (cudaStream_t)stream));
    h->launches += 1;
    return GPR_OK;
}

extern "C" int gpr_compute_reward_f64(gpr_handle* h, int batch, const double* achieved, const double* desired,
                                      const uint8_t* mover_collision, const uint8_t* wall_collision, float* reward,
                                      uint8_t* terminated, void* stream) {
    if (!h || !achieved || !desired) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    if (batch < 0) return fail(GPR_ERR_INVALID_ARG, "batch < 0");
    if (batch == 0) return GPR_OK;
    DeviceGuard g(h->device);
    CU(launch_compute_reward(h->cfg.env_kind, h->cfg.num_movers, batch, h->cfg.threshold_pos, achieved, desired, true, mover_collision,
                             wall_collision, reward, terminated, (cudaStream_t)stream));
    h->launches += 1;
    return GPR_OK;
}

extern "C" int gpr_episode_stats(gpr_handle* h, double* dst, int reset_after, void* stream) {
    if (!h || !dst) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    cudaStream_t s = (cudaStream_t)stream;
    CU(cudaMemcpyAsync(dst, h->stats, 6 * sizeof(double), cudaMemcpyDeviceToDevice, s));
    if (reset_after) CU(cudaMemsetAsync(h->stats, 0, 6 * sizeof(double), s));
    return GPR_OK;
}

extern "C" int gpr_invalidate_outputs(gpr_handle* h) {
    if (!h) return fail(GPR_ERR_INVALID_ARG, "handle is NULL");
    h->goal_dirty = true;
    h->goal_ptr = nullptr;
    h->host_goal_synced = nullptr;
    return GPR_OK;
}

extern "C" int gpr_debug_build(void) {
#ifdef GPR_DEBUG_BOUNDS
    return 1;
#else
    return 0;
#endif
}

extern "C" int gpr_debug_errors(gpr_handle* h, uint32_t* host_counts) {
    if (!h || !host_counts) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    CU(cudaDeviceSynchronize());
    CU(cudaMemcpy(host_counts, h->debug_errors, DBG_NUM_SLOTS * sizeof(uint32_t), cudaMemcpyDeviceToHost));
    // host-side invariants of the work list after a completed step: every slot handed back ("empty" pattern, planning) and
    // the consumers' cursor at or beyond the published count of the step that just ran (slot 6 / 7)
    host_counts[6] = host_counts[7] = 0;
    const size_t B = (size_t)h->cfg.num_envs;
    unsigned long long ctl[4];
    CU(cudaMemcpy(ctl, h->reset_count, sizeof(ctl), cudaMemcpyDeviceToHost));
    const int last = h->parity ^ 1;  // the buffer the last step used
    const uint32_t* cur = reinterpret_cast<const uint32_t*>(ctl + 2);
    const unsigned long long published = h->cfg.env_kind == GPR_ENV_PLANNING ? (ctl[last] & 0xffffffffull) : ctl[last];
    if (published > B) host_counts[7] += 1;
    if (h->cfg.autoreset_mode != GPR_AUTORESET_OFF || h->cfg.env_kind == GPR_ENV_PUSHING)
        if ((unsigned long long)cur[last] < published) host_counts[7] += 1;
    if (h->cfg.env_kind == GPR_ENV_PLANNING) {
        std::vector<unsigned long long> list(B);
        CU(cudaMemcpy(list.data(), h->reset_list, B * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < B; ++i) host_counts[6] += list[i] != ~0ull;
    }
    return GPR_OK;
}

extern "C" int gpr_reset_failures(gpr_handle* h, uint32_t* host_count) {
    if (!h || !host_count) return fail(GPR_ERR_INVALID_ARG, "NULL argument");
    DeviceGuard g(h->device);
    CU(cudaMemcpy(host_count, h->fail_count, sizeof(uint32_t), cudaMemcpyDeviceToHost));
    return GPR_OK;
}
