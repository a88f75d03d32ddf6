/*
 * gpr.h — C ABI of the B200-native batched simulator for GymPR's step path.
 *
 * The reference (ubi-coro/gymnasium-planar-robotics v1.1.0a2) is pure Python and has no FFI for this path: its "operator
 * API" is gymnasium.Env (reset/step/compute_reward/...).  The entry points below are what a Python binding of that API
 * needs; each cites the reference interface it replaces (paths relative to gymnasium_planar_robotics/):
 *
 *   gpr_create         <- BasicPlanarRoboticsEnv.__init__            envs/basic_envs.py:162-289
 *                         BenchmarkPlanningEnv.__init__              envs/planning/benchmark_planning_env.py:165-291
 *                         BenchmarkPushingEnv.__init__               envs/manipulation/benchmark_pushing_env.py:154-288
 *   gpr_reset          <- BasicPlanarRoboticsSingleAgentEnv.reset    envs/basic_envs.py:1770-1833
 *                         _reset_callback                            planning:355-418, pushing:373-417
 *   gpr_step           <- BasicPlanarRoboticsSingleAgentEnv.step     envs/basic_envs.py:1835-1950
 *                         (+ gymnasium TimeLimit(50)                 __init__.py:25-38)
 *   gpr_step_host      <- same call made with host (NumPy) buffers, as a user of the reference makes it
 *   gpr_compute_reward <- compute_reward / compute_terminated        planning:459-534, pushing:457-527   (HER relabelling)
 *   gpr_get_state /
 *   gpr_set_state      <- MjData.qpos/qvel/act/qacc access           utils/mujoco_utils.py:23-190 (parity + checkpoint)
 *   gpr_get_seed /
 *   gpr_set_seed       <- np_random / rng_noise of reset(seed)       envs/basic_envs.py:1789-1791 (checkpoint)
 *   gpr_destroy        <- Env.close                                  planning:604-608
 *
 * Conventions
 *   - plain C, POD structs, raw pointers; no C++/torch types cross this boundary.
 *   - every function returns GPR_OK (0) or a negative gpr_status; nothing throws; gpr_last_error() gives the text
 *     (thread-local).  Device code never asserts: an out-of-grid mover is a wall collision (reference: AssertionError,
 *     basic_envs.py:514-517).
 *   - the caller owns all I/O buffers (device pointers unless the function name ends in _host); the handle owns only its
 *     structure-of-arrays state.  All work is enqueued on the caller's stream (a cudaStream_t passed as void*); no hidden
 *     synchronisation except in the *_host calls, which return when the host buffers are filled.
 *   - one handle <-> one device; a handle is not thread-safe.
 *   - state and arithmetic are IEEE float64 (the reference's dtype) evaluated in the reference's operation order without
 *     FMA contraction, so collision/termination flags are bit-identical to the float64 oracle; observations, goals and
 *     rewards are delivered as float32 (the float64 value rounded once on store); actions are float32.
 */
#ifndef GPR_H_
#define GPR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPR_ABI_VERSION 3

#if defined(__GNUC__)
#define GPR_API __attribute__((visibility("default")))
#else
#define GPR_API
#endif

#define GPR_MAX_MOVERS 32   /* one lane per mover, one lane group (<= one warp) per environment */
#define GPR_MAX_TILES_1D 32 /* tiles per axis */
#define GPR_MAX_OBSTACLES 8 /* static obstacles per layout (planning env) */

enum gpr_env_kind { GPR_ENV_PLANNING = 0, GPR_ENV_PUSHING = 1 };
enum gpr_collision_shape { GPR_SHAPE_CIRCLE = 0, GPR_SHAPE_BOX = 1 };
enum gpr_autoreset_mode {
    GPR_AUTORESET_OFF = 0,       /* like the bare reference env: the caller resets; counters keep running */
    GPR_AUTORESET_SAME_STEP = 1, /* finished envs are re-sampled inside the same kernel; terminal obs go to final_* */
    GPR_AUTORESET_NEXT_STEP = 2  /* gymnasium>=1.0 vector default: the step after `done` performs the reset instead */
};

typedef enum gpr_status {
    GPR_OK = 0,
    GPR_ERR_INVALID_ARG = -1,
    GPR_ERR_CUDA = -2,
    GPR_ERR_NO_DEVICE = -3,
    GPR_ERR_OUT_OF_MEMORY = -4,
    GPR_ERR_UNSUPPORTED = -5,
    GPR_ERR_ABI_MISMATCH = -6
} gpr_status;

/*
 * Everything the step path needs, packed once by the host mirror (the Python env class) from the reference's constructor
 * kwargs.  Derived numbers are computed on the host in float64 WITH THE REFERENCE'S EXPRESSIONS so thresholds are
 * bit-identical (citations per field).
 */
typedef struct gpr_config {
    uint32_t struct_bytes; /* sizeof(gpr_config) as seen by the caller; gpr_create rejects a mismatch */
    uint32_t abi_version;  /* GPR_ABI_VERSION */
    int32_t env_kind;      /* gpr_env_kind */
    int32_t num_envs;      /* environments held by this handle (this rank's shard) */
    int64_t env_index_base; /* global index of local env 0; part of the RNG key, so results do not depend on sharding */
    uint64_t seed;          /* base RNG key; gpr_reset may replace it */

    /* --- tile layout (basic_envs.py:194-221, 1292-1339) --- */
    int32_t num_tiles_x, num_tiles_y;
    uint8_t layout[GPR_MAX_TILES_1D * GPR_MAX_TILES_1D]; /* layout[ix * num_tiles_y + iy] in {0,1} */
    double tile_half[2];                                 /* tile_size[:2] (half sizes), default 0.12 */
    double tile_cx[GPR_MAX_TILES_1D];                    /* np.linspace centres along x (basic_envs.py:1300-1303) */
    double tile_cy[GPR_MAX_TILES_1D];

    /* --- collision shapes (basic_envs.py:257-264) --- */
    int32_t c_shape;          /* gpr_collision_shape */
    int32_t reference_quirks; /* 1: reproduce the (P,)x(P,1) broadcast of basic_envs.py:409 for per-mover radii */
    /* [safety][mover][xy]; circle uses [.][.][0] only.
       c_wall  = c_size + offset_wall + safety*offset   (basic_envs.py:487)
       c_mover = c_size + safety*offset                 (basic_envs.py:390)                                   */
    double c_wall[2][GPR_MAX_MOVERS][2];
    double c_mover[2][GPR_MAX_MOVERS][2];

    /* --- dynamics (planning:213-218, 420-450; basic_envs.py:283, 1740) --- */
    int32_t num_movers;
    int32_t learn_jerk;
    int32_t num_cycles; /* control cycles per env-step, default 40 */
    int32_t max_episode_steps; /* TimeLimit, 50 (__init__.py:28,37); <=0 disables truncation */
    double cycle_time;  /* MuJoCo opt.timestep, 0.001 */
    double v_max, a_max, j_max;

    /* --- task (planning:220, 262-274; pushing:218, 250-288) --- */
    double threshold_pos;
    double min_xy_pos[2], max_xy_pos[2]; /* mover spawn box */
    double min_goal_dist;                /* planning: strict '<' rejects (planning:410) */
    double std_noise[3];                 /* sigma for position, velocity, acceleration (basic_envs.py:184-192) */
    int32_t autoreset_mode;              /* gpr_autoreset_mode */
    int32_t max_reset_attempts;          /* cap on the rejection-sampling loops (reference: unbounded) */

    /* --- pushing only (pushing:172-178, 250-288, 332-342; utils/impedance_control.py:28-55) --- */
    double object_min_xy_pos[2], object_max_xy_pos[2];
    double min_mo_dist;     /* strict '>' accepts (pushing:404) */
    double object_noise_xy; /* 1e-5 (pushing:178) */
    double object_half_xy;  /* 0.035 */
    double object_mass;     /* 0.01 */
    double object_damping;  /* free-joint damping 0.01 on every DOF (pushing:337) */
    double mover_half[2];   /* physical mover half sizes (0.0775, 0.0775) */
    double mover_mass;      /* 1.24 */
    double imp_k_rot;       /* rotational stiffness 0.1 */
    double gravity;         /* 9.81 */
    double friction;        /* MuJoCo default geom friction 1.0 */
    double solref[2];       /* MuJoCo default (0.02, 1) */
    double solimp[5];       /* MuJoCo default (0.9, 0.95, 0.001, 0.5, 2) */
    int32_t contact_iterations; /* projected Gauss-Seidel sweeps of the planar contact solve */
    int32_t output_flags; /* GPR_OUT_* bits */

    /* --- static obstacles, planning env only: the typed form of the reference's extension point
     * `_check_for_other_collisions_callback` (basic_envs.py:1976-1986; called at basic_envs.py:1807 and 1903).  The
     * reference's hook returns False in both benchmark envs; a custom env overrides it in Python.  Here an obstacle is a
     * fixed shape of the env's own collision kind (c_shape): a circle (centre, radius) or an axis-aligned box (centre,
     * half sizes), the same in every environment.  A mover touches an obstacle under the rules of the mover-mover check
     * (basic_envs.py:390-424: circle  ||p - o|| <= r_mover + r_obstacle;  box  geom.check_rectangles_intersect, or the
     * mover's centre inside the obstacle), evaluated on the noisy qpos of the WALL check (static geometry is checked
     * together: no additional random draws).  The flag is reported as `other_collision`, ends the 40-cycle loop like the
     * other two (basic_envs.py:1904) and counts as a collision in reward / termination; starts and goals are re-sampled
     * until they clear every obstacle by the safety offset.  num_obstacles = 0 reproduces the reference exactly. */
    int32_t num_obstacles;
    int32_t contact_warm_start; /* pushing: start the contact solve from the previous substep's forces (include/gpr_push_physics.h) */
    double obstacle_xy[GPR_MAX_OBSTACLES][2];
    double obstacle_size[GPR_MAX_OBSTACLES][2]; /* circle: radius in [0]; box: half sizes */
    /* --- typed "extra bodies" (SURVEY.md §8f-3): what a custom env of the reference adds to the MuJoCo model through
     * `custom_model_xml_strings` (basic_envs.py:134-155; docs/make_own_env.rst) and then checks in its
     * `_check_for_other_collisions_callback` — here without XML, as kinematic planar bodies: each obstacle may carry a
     * PRESCRIBED constant velocity.  Its centre at the check of control cycle c (0-based) of an episode's env-step s
     * (0-based) is  obstacle_xy + obstacle_vel * ((s * num_cycles + c + 1) * cycle_time)  — product and sum each rounded
     * once in float64 — i.e. the body has moved for as long as the movers have been integrated; at reset() (and for the
     * start / goal sampling) it is at obstacle_xy.  All zeros = static obstacles = the behaviour of ABI version 2. */
    double obstacle_vel[GPR_MAX_OBSTACLES][2];
} gpr_config;

/* output_flags bit: a step writes `desired_goal` rows only for environments whose goal changed in that call (they were
 * reset); the other rows keep what an earlier call wrote.  For callers that pass the SAME desired_goal buffer to every
 * call (the Python env classes do): it takes a quarter off the result traffic, which is what bounds *_host calls.
 * gpr_reset always writes the rows of the environments it resets; after gpr_set_state with a goal, the next step writes
 * every row once.
 * The library recognises "the same buffer" by its ADDRESS: a caller that frees and re-allocates its desired_goal buffer (a
 * caching allocator may well hand back the same address) must call gpr_invalidate_outputs() before the next step, which
 * then writes every row once.  Callers that cannot guarantee a persistent buffer should leave this flag clear. */
#define GPR_OUT_GOAL_ON_CHANGE 1
/* output_flags bit: observation, achieved_goal, desired_goal and the final_* arrays are float64 (`double*` behind the
 * `float*` fields of gpr_outputs): the reference's dtype.  The kernels compute in float64 and decide goal-reached / reward /
 * terminated on the float64 values; float32 outputs are those values rounded once, so a caller that RE-derives the reward
 * from float32 goals (HER: compute_reward(achieved_goal, desired_goal, info)) can disagree with the step's own reward when
 * a distance lies within float32 rounding (~3e-8 m) of threshold_pos — about once per 1e6 mover-steps.  With this flag the
 * outputs are the very values the step decided on and gpr_compute_reward_f64 reproduces its reward and termination
 * exactly, as the reference guarantees (its spaces are float64).  The single-env classes of the Python binding set it;
 * the vector classes default to float32 (half the result traffic).  reward stays float32 (-50, +50, -k: exact). */
#define GPR_OUT_FLOAT64 2

/* Per-step results. Device pointers (host pointers for *_host calls), caller-owned, row-major; NULL = do not write.
 * planning: obs_dim = 2*N*(1+learn_jerk) (planning:242-254), goal_dim = 2*N
 * pushing : obs_dim = 2*(2+learn_jerk)   (pushing:232-247),  goal_dim = 2                                           */
typedef struct gpr_outputs {
    float* observation;   /* [num_envs, obs_dim]  */
    float* achieved_goal; /* [num_envs, goal_dim] */
    float* desired_goal;  /* [num_envs, goal_dim] */
    float* reward;        /* [num_envs] */
    uint8_t* terminated;  /* [num_envs] compute_terminated (planning:459-479, pushing:457-476) */
    uint8_t* truncated;   /* [num_envs] TimeLimit: elapsed_steps >= max_episode_steps */
    uint8_t* is_success;  /* [num_envs] info['is_success'] (planning:596-601) */
    uint8_t* mover_collision;
    uint8_t* wall_collision;
    /* SAME_STEP autoreset only: terminal observation of the envs that finished in this step (rows of other envs are
       left untouched).  observation/achieved_goal/desired_goal above then hold the first observation of the new episode. */
    float* final_observation;
    float* final_achieved_goal;
    float* final_desired_goal;
    uint8_t* other_collision; /* [num_envs] a mover touches a static obstacle (see gpr_config.num_obstacles); all 0 without obstacles */
    /* COMPACT terminal observations — gpr_step_host only, optional (NULL = the dense form above).  Like gymnasium's vector
       envs, which report final observations only for the sub-envs that finished, a host caller may take them as a list:
       when the call uses the copy-engine route with compact transport (several ranks on one host, see gpr_step_host) and
       final_index is non-NULL, row s of final_observation / final_achieved_goal / final_desired_goal belongs to env
       final_index[s] for s < *final_count (order unspecified) and NO row is scattered on the host.  Any other route fills
       final_* densely as usual and sets *final_count = UINT32_MAX.  Both arrays are caller-owned: final_index
       [num_envs] int32, final_count [1] uint32. */
    int32_t* final_index;
    uint32_t* final_count;
} gpr_outputs;

/* Structure-of-arrays state, float64. Device pointers, caller-owned copies; NULL = skip that field.
 * planning: pos/vel/acc/goal are [num_envs, N, 2]; `acc` is MuJoCo's qacc == the jerk integrator state `act`.
 * pushing : pos/vel/acc are [num_envs, 1, 2]; goal is [num_envs, 2]; the extra fields below are used.                */
typedef struct gpr_state {
    double* pos;
    double* vel;
    double* acc;
    double* goal;
    int32_t* elapsed_steps; /* [num_envs] TimeLimit counter */
    uint32_t* rng_counter;  /* [num_envs] number of step/reset events consumed so far (Philox counter word) */
    /* pushing only */
    double* act;         /* [num_envs, 2] jerk-mode integrator state (differs from qacc under contact) */
    double* mover_rot;   /* [num_envs, 3] cos(yaw), sin(yaw), yaw rate (orientation kept on the unit circle) */
    double* object_pos;  /* [num_envs, 4] x, y, cos(yaw), sin(yaw) */
    double* object_vel;  /* [num_envs, 3] vx, vy, yaw rate */
    /* bookkeeping needed to resume an interrupted run exactly (checkpoint / restore) */
    uint8_t* needs_reset;   /* [num_envs] NEXT_STEP auto-reset: the env finished in the previous step */
    float* episode_return;  /* [num_envs] return accumulated so far in the running episode (episode statistics) */
    float* contact_warm;    /* pushing only: [num_envs, 13] warm-start state of the contact solve (GPR_PUSH_WARM) */
} gpr_state;

typedef struct gpr_handle gpr_handle;

/* sizeof(gpr_config) / ABI version as compiled into the library (binding self-check). */
GPR_API uint32_t gpr_config_bytes(void);
GPR_API uint32_t gpr_abi_version(void);

/* Text of the last error on the calling thread ("" if none). */
GPR_API const char* gpr_last_error(void);

/* Allocate the SoA state for cfg->num_envs environments on `device`. State is undefined until gpr_reset. */
GPR_API int gpr_create(const gpr_config* cfg, int device, gpr_handle** out_handle);
GPR_API void gpr_destroy(gpr_handle* h);

/* Dimensions implied by the config (so bindings need not duplicate the formulas). */
GPR_API int gpr_obs_dim(const gpr_handle* h);
GPR_API int gpr_goal_dim(const gpr_handle* h);
GPR_API int gpr_action_dim(const gpr_handle* h);

/*
 * Start new episodes.
 *   reset_mask   : [num_envs] device bytes, non-zero = reset that env; NULL = all.
 *   seed         : if reseed != 0 the RNG key becomes `seed` and all RNG counters restart at 0 (reference: reset(seed=..)).
 *   inject_start : NULL -> rejection-sample like planning:369-385 / pushing:386-404; else device float64
 *                  [num_envs, N, 2] mover start positions taken as given (parity tests; reference: reload_model(...)).
 *   inject_goal  : same for goals [num_envs, goal_dim/2, 2] (pushing: object goal).
 *   inject_object: pushing only, [num_envs, 2] object start position; NULL -> sampled.
 *   out          : observation / achieved / desired / is_success / mover_collision / wall_collision are written for the
 *                  reset envs (reset() returns (obs, info), basic_envs.py:1807-1833); reward/terminated/truncated untouched.
 */
GPR_API int gpr_reset(gpr_handle* h, const uint8_t* reset_mask, int reseed, uint64_t seed, const double* inject_start,
              const double* inject_goal, const double* inject_object, const gpr_outputs* out, void* stream);

/* One env-step for every environment: action clip, num_cycles x {limit control, integrate, wall + mover checks, break on
 * collision}, observation, info, reward, terminated, truncated, optional auto-reset.  action: device float32
 * [num_envs, action_dim]. */
GPR_API int gpr_step(gpr_handle* h, const float* action, const gpr_outputs* out, void* stream);

/* Same step, called the way a user of the reference calls it: HOST buffers in, HOST buffers out; returns after the host
 * buffers are valid.  `host_out` holds host pointers.  Page-locked buffers (cudaHostAlloc / cudaHostRegister / torch
 * pinned tensors) are read and written by the kernels IN PLACE through their device alias (zero-copy: result stores cross
 * PCIe while the rest of the grid still computes); pageable buffers are staged through the handle's pinned mirror.
 * Environment variable GPR_HOST_IO (read once per process): `zerocopy` as above; `dma` routes the results of page-locked
 * buffers through device staging and the copy engine instead — slower on an otherwise idle host (B200: 192M -> 119M
 * env-steps/s at 65,536 envs), faster when many GPUs write into one host at once (8 ranks: 436M -> 497M); unset = `auto`:
 * dma when the process is one of several ranks on the host (LOCAL_WORLD_SIZE / WORLD_SIZE > 1), zero-copy otherwise. */
GPR_API int gpr_step_host(gpr_handle* h, const float* host_action, const gpr_outputs* host_out);
GPR_API int gpr_reset_host(gpr_handle* h, int reseed, uint64_t seed, const gpr_outputs* host_out);

/* Copy state out of / into the handle (device pointers, float64). */
GPR_API int gpr_get_state(gpr_handle* h, const gpr_state* dst, void* stream);
GPR_API int gpr_set_state(gpr_handle* h, const gpr_state* src, void* stream);

/* The RNG key in use (gpr_reset with reseed != 0 replaces the one of the config).  gpr_set_seed changes the key WITHOUT
 * touching the per-env event counters: together with gpr_set_state it restores a checkpoint exactly. */
GPR_API int gpr_get_seed(const gpr_handle* h, uint64_t* seed);
GPR_API int gpr_set_seed(gpr_handle* h, uint64_t seed);

/* HER relabelling: batched compute_reward / compute_terminated on device.
 *   achieved, desired : float32 [batch, goal_dim]; mover_collision, wall_collision: [batch] bytes (NULL = all false)
 *   reward : float32 [batch] or NULL; terminated : bytes [batch] or NULL                                             */
GPR_API int gpr_compute_reward(gpr_handle* h, int batch, const float* achieved, const float* desired,
                       const uint8_t* mover_collision, const uint8_t* wall_collision, float* reward, uint8_t* terminated,
                       void* stream);

/* The same for float64 goals (GPR_OUT_FLOAT64 outputs): achieved, desired are double [batch, goal_dim]. */
GPR_API int gpr_compute_reward_f64(gpr_handle* h, int batch, const double* achieved, const double* desired,
                           const uint8_t* mover_collision, const uint8_t* wall_collision, float* reward, uint8_t* terminated,
                           void* stream);

/* Episode statistics accumulated on device since the last call with reset_after != 0 (6 float64 values):
 * [episodes finished, sum of returns, sum of lengths, successes, mover collisions, wall collisions].
 * `dst` is a device pointer; the caller may all-reduce it over ranks (NCCL) — nothing on the step path communicates. */
GPR_API int gpr_episode_stats(gpr_handle* h, double* dst, int reset_after, void* stream);

/* Number of resets whose rejection-sampling loop hit max_reset_attempts since creation (synchronous D2H read). */
GPR_API int gpr_reset_failures(gpr_handle* h, uint32_t* host_count);

/* Per-kernel device times of gpr_step (measurement aid, off by default).  enable != 0: every following gpr_step brackets
 * its kernels with CUDA events on the caller's stream.  host_ms (may be NULL) receives, after synchronising those events,
 * [0] total ms of the step kernel, [1] total ms of the auto-reset kernel, [2] number of gpr_step calls covered, and the
 * accumulation restarts.  enable == 0 stops recording. */
GPR_API int gpr_kernel_times(gpr_handle* h, int enable, double* host_ms);

/* GPR_OUT_GOAL_ON_CHANGE: forget which desired_goal buffer has been written — the next gpr_step writes every row. */
GPR_API int gpr_invalidate_outputs(gpr_handle* h);

/* Memory-safety aid (compute-sanitizer is not available on every pool): a library built with -DGPR_DEBUG_BOUNDS
 * (csrc/Makefile target `debug` -> libgpr_b200_dbg.so; gpr_debug_build() returns 1) range-checks every index its kernels
 * derive for global memory and counts violations.  gpr_debug_errors synchronises the device and fills host_counts[8]:
 * [0] lane (env, mover) index, [1] work-list slot, [2] work-list entry, [3] consumer claim beyond the published count,
 * [4] output row, [5] shared-memory staging index (all always 0 in a release build), [6] work-list slots not handed back
 * after the last step, [7] work-list counters inconsistent (checked on the host in every build). */
GPR_API int gpr_debug_build(void);
GPR_API int gpr_debug_errors(gpr_handle* h, uint32_t* host_counts);

/* Number of kernels this library has launched on behalf of the handle (bench "gpu_launches" evidence). */
GPR_API uint64_t gpr_launch_count(const gpr_handle* h);

#ifdef __cplusplus
}
#endif
#endif /* GPR_H_ */
