/*
 * tde_b200.h — C ABI of libtde_b200.so, the B200 (sm_100a) implementation of
 * TorchDriveEnv's per-timestep simulation hot path.
 *
 * Every entry point replaces one call the reference makes on its
 * SimulatorInterface-level surface (citations are file:line under the
 * reference checkout, torchdriveenv/gym_env.py unless said otherwise):
 *
 *   tde_step                     WaypointSuiteEnv.step :369-389 + GymEnv.step :115-120
 *                                  (simulator.step :117, get_obs :122-124, get_reward :396-411,
 *                                   is_terminated :413-417, is_truncated :134-135, get_info :419-437,
 *                                   check_reach_target/advance :378-383,391-394)
 *   tde_reset                    WaypointSuiteEnv.reset :319-349 + set_start_pos :351-367
 *                                  + build_simulator's state initialisation :192-198,241-247,269-283
 *   tde_render                   simulator.render_egocentric() :123,154
 *   tde_step_stacked,
 *   tde_render_stacked           the same with VecFrameStack (examples/rl_training.py:160) fused into the store
 *   tde_step_terminal            the same, keeping info["terminal_observation"] of SB3's VecEnv (the caller's contract)
 *   tde_step_rollout,
 *   tde_step_rollout_scatter     the same, writing the next slot of a rollout buffer (collect_rollouts of the
 *                                  trainer that drives the env, examples/rl_training.py:178-181,200)
 *   tde_get_state/tde_set_state  simulator.get_state() :127,371,392-393,397-399,420-423 / set_state :247
 *   tde_compute_infractions,
 *   tde_get_infractions          simulator.compute_offroad() :142,415,427; compute_collision() :143,415,428;
 *                                  compute_traffic_lights_violations() :144,415,429; compute_wrong_way()
 *   tde_kinematics               KinematicBicycle.step via simulator.step :117 (+ IAIWrapper replay :275-294)
 *   tde_collision_boxes,
 *   tde_offroad_boxes            the same metrics on caller-provided boxes (micro-bench config C4)
 *   tde_upload_scenarios         the constructor-side tensors of build_simulator :179-300
 *                                  (road_mesh :184,260; traffic_controls :187-189; waypoints :252-257;
 *                                   agent_states/attributes :241-247,261; replay_states/replay_mask :275-283)
 *   tde_clone                    simulator.copy() :110
 *   tde_get_episode_stats        the per-episode quantities EvalNTimestepsCallback aggregates
 *                                  (examples/rl_training.py:99-108)
 *
 * Conventions
 *   - Plain C, no torch types.  All *_dev pointers are CUDA device pointers owned by the caller;
 *     *_host pointers are host memory.  `stream` is a cudaStream_t passed as void*.
 *   - Every function returns 0 (TDE_OK) or a negative TDE_E_* code; the message is available through
 *     tde_last_error().  Nothing throws or exits across this boundary.
 *   - tde_step / tde_render / tde_kinematics / ... never allocate, never synchronise with the host and
 *     only enqueue work on `stream`, so they can be captured into a CUDA graph.  The exceptions say so at
 *     their declaration: tde_step_host (allocates device and pinned staging buffers on first use, runs its device-to-host
 *     copies on an internal second stream, expands the frames on a pool of host threads it owns and synchronises
 *     `stream` before it returns), tde_render_view (may grow
 *     its scratch), tde_get_episode_stats (synchronises), tde_create / tde_upload_scenarios / tde_clone (allocate).
 *     A step without observations (obs_dev == NULL) is launched with programmatic stream serialization unless `stream`
 *     is capturing: it may be scheduled while the previous kernel on the stream drains and waits for it on the device.
 *   - A handle is bound to one GPU and is not thread-safe.  Every call runs on the handle's GPU and restores the
 *     caller's current CUDA device before it returns.
 *   - The seed belongs to the handle: tde_reset with env_mask_dev == NULL sets it, a masked reset keeps it.
 */
#ifndef TDE_B200_H
#define TDE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TDE_VERSION 100 /* 0.1.0 */

/* error codes */
#define TDE_OK 0
#define TDE_E_INVAL (-1) /* bad argument */
#define TDE_E_CUDA (-2)  /* CUDA runtime error */
#define TDE_E_SHAPE (-3) /* size/shape mismatch or capacity exceeded */
#define TDE_E_ARCH (-4)  /* device is not sm_100 */
#define TDE_E_STATE (-5) /* call order (e.g. step before upload/reset) */

/* limits of this build */
#define TDE_MAX_AGENTS 64
#define TDE_OBS_H 64
#define TDE_OBS_W 64
#define TDE_OBS_C 3
#define TDE_MAX_STOPLINES 32
#define TDE_INFO_STRIDE 16

/* info[] columns, one row of TDE_INFO_STRIDE floats per env (get_info :419-437) */
#define TDE_INFO_OFFROAD 0
#define TDE_INFO_COLLISION 1
#define TDE_INFO_TL_VIOLATION 2
#define TDE_INFO_IS_SUCCESS 3
#define TDE_INFO_REACHED_WAYPOINT_NUM 4
#define TDE_INFO_PSI_SMOOTHNESS 5
#define TDE_INFO_PSI_REWARD 6
#define TDE_INFO_DIST_REWARD 7
#define TDE_INFO_SPEED_SMOOTHNESS 8
#define TDE_INFO_WRONG_WAY 9
#define TDE_INFO_EPISODE_RETURN 10
#define TDE_INFO_EPISODE_LENGTH 11
#define TDE_INFO_SCENARIO 12
#define TDE_INFO_DID_RESET 13

/* per-agent infraction cache columns (float4 per agent) */
#define TDE_INFR_COLLISION 0
#define TDE_INFR_OFFROAD 1
#define TDE_INFR_TL_VIOLATION 2
#define TDE_INFR_WRONG_WAY 3

/* render classes = painter's levels, higher paints over lower */
#define TDE_CLS_BACKGROUND 0
#define TDE_CLS_ROAD 1
#define TDE_CLS_LANE_MARKING 2
#define TDE_CLS_TL_GREEN 3
#define TDE_CLS_TL_YELLOW 4
#define TDE_CLS_TL_RED 5
#define TDE_CLS_WAYPOINT 6
#define TDE_CLS_VEHICLE 7
#define TDE_CLS_EGO 8
#define TDE_CLS_DIRECTION 9
#define TDE_CLS_EGO_DIRECTION 10
#define TDE_NUM_CLASSES 11

/* traffic-light states in the schedule table */
#define TDE_LIGHT_GREEN 0
#define TDE_LIGHT_YELLOW 1
#define TDE_LIGHT_RED 2

/* phase bits for tde_step_phases (tde_step == all of them) */
#define TDE_PH_KINEMATICS 1
#define TDE_PH_INFRACTIONS 2
#define TDE_PH_REWARD 4
#define TDE_PH_RENDER 8
#define TDE_PH_ALL 15

/* episode statistics vector (doubles), reduced across GPUs by the host side */
#define TDE_STAT_EPISODES 0
#define TDE_STAT_RETURN_SUM 1
#define TDE_STAT_LENGTH_SUM 2
#define TDE_STAT_OFFROAD 3
#define TDE_STAT_COLLISION 4
#define TDE_STAT_TL_VIOLATION 5
#define TDE_STAT_SUCCESS 6
#define TDE_STAT_REACHED_WAYPOINTS 7
#define TDE_STAT_STEPS 8
#define TDE_NUM_STATS 16

typedef struct tde_handle tde_handle;

/* EnvConfig :34-54 plus the TorchDriveConfig/RendererConfig values the env relies on. */
typedef struct tde_config {
    int32_t num_envs;              /* E, envs owned by this handle (this GPU's shard) */
    int32_t max_agents;            /* A, agent slots per env, slot 0 = ego; 1..TDE_MAX_AGENTS */
    int64_t env_index_offset;      /* global index of local env 0 (keys the reset RNG, so a sharded run
                                      reproduces the unsharded one) */
    int32_t max_environment_steps; /* :37  default 200 */
    int32_t terminated_at_infraction; /* :44 default 1 */
    int32_t left_handed_coordinates;  /* :46-49 default 1 */
    int32_t auto_reset;            /* 1: finished envs are re-initialised inside tde_step (VecEnv semantics) */
    int32_t randomize_ego_attributes; /* 1: ego length/width/lr ~ U as in :194-196, else scenario table */
    int32_t device;                /* CUDA ordinal */
    float dt;                      /* 0.1 (:432) */
    float waypoint_bonus;          /* :39 100 */
    float heading_penalty;         /* :40 25 */
    float distance_bonus;          /* :41 1 */
    float distance_cutoff;         /* :42 0.5 */
    float reach_radius;            /* :394 3 */
    float offroad_threshold;       /* TorchDriveConfig default 0.5 [EXT-RECALLED] */
    float tl_rear_factor;          /* fraction of the agent box (at its rear) tested against red stop lines */
    float fov;                     /* metres covered by the 64 px birdview, default 35 */
    float start_speed_max;         /* :358 10 */
    float start_heading_sigma;     /* :361 0.1 */
    int32_t stage_map_tables;      /* the physics kernel can copy the per-map tables (lane-mesh triangle records, stop lines,
                                      light schedule, per-cell summary) into shared memory once per CTA with bulk-async
                                      copies (cp.async.bulk + mbarrier) when they fit, instead of reading them through L1.
                                      0 (default): it does for handles of at most 8 envs per SM (small batches start every
                                      launch with a cold L1), 1: always, 2: never.  TDE_PHYS_STAGE=0|1 overrides it. */
    int32_t host_obs_rgb;          /* tde_step_host only.  0 (default): the frames cross PCIe as the 4-bit class image (2 KB
                                      per env) and host threads (TDE_HOST_THREADS; default: the CPUs of the process /
                                      LOCAL_WORLD_SIZE, at most 16) expand them to the caller's RGB planes with the
                                      palette (AVX-512 / AVX2 / scalar, whatever the CPU has); 1: the RGB planes (12 KB per env) cross PCIe and no host thread is
                                      started.  Same bytes in the caller's buffer either way.  The environment variable
                                      TDE_HOST_OBS=classes|rgb overrides it. */
    int32_t reserved[6];
} tde_config;

/*
 * Scenario tables (host pointers, copied by tde_upload_scenarios).
 *
 * A "map" is static geometry shared by scenarios: the road (lane) mesh with a lane direction per
 * triangle, lane-marking triangles, stop lines and their light schedule.  A "scenario" is a waypoint
 * polyline on a map plus the agent table (slot 0 = ego) and the optional NPC log-replay.
 */
typedef struct tde_scenario_set {
    int32_t num_maps;
    /* road mesh: map m owns triangles [map_tri_offset[m], map_tri_offset[m+1]);
       8 floats per triangle: ax ay bx by cx cy lane_dir_cos lane_dir_sin */
    const int32_t* map_tri_offset;
    const float* road_tris;
    /* lane markings: 6 floats per triangle */
    const int32_t* map_mark_offset;
    const float* mark_tris;
    /* stop lines: 5 floats per line: x y length width psi */
    const int32_t* map_stop_offset;
    const float* stoplines;
    /* light schedule: map m has period P=map_light_period[m] (>=1) rows of L(m) uint8 states starting
       at light_states[map_light_offset[m]], row-major [t][l] */
    const int32_t* map_light_period;
    const int32_t* map_light_offset;
    const uint8_t* light_states;

    int32_t num_scenarios;
    const int32_t* scen_map;           /* [Ns] */
    const int32_t* scen_wp_offset;     /* [Ns+1] */
    const float* waypoints;            /* 2 floats per waypoint */
    const float* scen_start_heading;   /* [Ns] lane direction at the start segment (find_lanelet_directions :359) */
    const int32_t* scen_num_agents;    /* [Ns] present agents incl. ego, <= max_agents */
    const float* agent_init;           /* [Ns][A][4] x y psi v (slot 0 ignored: ego start is sampled) */
    const float* agent_attr;           /* [Ns][A][3] length width rear_axis_offset */
    const int32_t* scen_replay_T;      /* [Ns] replay horizon (0 = none) */
    const int32_t* scen_replay_offset; /* [Ns+1] in time rows; row r holds A states / A mask bytes */
    const float* replay_states;        /* [rows][A][4] */
    const uint8_t* replay_mask;        /* [rows][A] */
} tde_scenario_set;

int tde_version(void);
/* message of the last failure on this handle (or of the last failed tde_create when h == NULL) */
const char* tde_last_error(const tde_handle* h);

int tde_create(const tde_config* cfg, tde_handle** out);
int tde_destroy(tde_handle* h);
/* default-initialise a config with the reference defaults */
int tde_default_config(tde_config* cfg);

int tde_upload_scenarios(tde_handle* h, const tde_scenario_set* set);
/* per-env scenario choice at reset is uniform over [lo[e], hi[e]); default [0, num_scenarios) */
int tde_set_env_scenario_range(tde_handle* h, const int32_t* lo_host, const int32_t* hi_host);
/* RGB palette, TDE_NUM_CLASSES x 3 bytes */
int tde_set_palette(tde_handle* h, const uint8_t* rgb_host);

/* env_mask_dev: E bytes (non-zero = reset) or NULL = all envs */
int tde_reset(tde_handle* h, const uint8_t* env_mask_dev, uint64_t seed, void* stream);

/* One lockstep env step for all E envs.
   actions_dev    float[E][2] (acceleration, steering) for the ego of each env
   obs_dev        uint8[E][3][64][64]      (may be NULL: no render)
   reward_dev     float[E]
   terminated_dev uint8[E]
   truncated_dev  uint8[E]
   info_dev       float[E][TDE_INFO_STRIDE] */
int tde_step(tde_handle* h, const float* actions_dev, uint8_t* obs_dev, float* reward_dev,
             uint8_t* terminated_dev, uint8_t* truncated_dev, float* info_dev, void* stream);
/* same with a TDE_PH_* mask: granular entry points for parity tests and micro-benchmarks */
int tde_step_phases(tde_handle* h, int32_t phases, const float* actions_dev, uint8_t* obs_dev,
                    float* reward_dev, uint8_t* terminated_dev, uint8_t* truncated_dev,
                    float* info_dev, void* stream);
/* host-buffer variant: copies actions H2D, steps, copies results D2H, synchronises `stream`.  NOT graph-capturable:
   it allocates device staging buffers on first use and, from 4,096 envs on, steps the envs in chunks whose frames
   travel back on an internal second stream while the next chunk is computed. */
int tde_step_host(tde_handle* h, const float* actions_host, uint8_t* obs_host, float* reward_host,
                  uint8_t* terminated_host, uint8_t* truncated_host, float* info_host, void* stream);

/* VecFrameStack(n_stack, channels_order="first") fused into the observation store
   (examples/rl_training.py:160): stack_dev is uint8[E][3*n_stack][64][64], oldest frame first.  The
   render kernel moves frames 1..n-1 of each env down by one slot while it writes the new frame into
   the last slot; an env that was reset since its previous frame (tde_reset, or the in-kernel
   auto-reset) has its older slots zeroed instead, as VecFrameStack does.  n_stack in 1..8. */
int tde_step_stacked(tde_handle* h, const float* actions_dev, uint8_t* stack_dev, int32_t n_stack, float* reward_dev,
                     uint8_t* terminated_dev, uint8_t* truncated_dev, float* info_dev, void* stream);
int tde_render_stacked(tde_handle* h, uint8_t* stack_dev, int32_t n_stack, void* stream);
/* Rollout collection (the on-policy loop that drives the env, examples/rl_training.py:159-160,178-181:
   VecFrameStack output stored per step into the rollout buffer): the same step, but the older frames
   are taken from stack_prev_dev (the observation slot of step t) while the shifted stack with the new
   frame is written to stack_next_dev (slot t + 1), so the buffer is filled without a copy pass.  Both
   are uint8[E][3*n_stack][64][64]; they must be the same buffer (= tde_step_stacked) or not overlap.
   reward / terminated / truncated / info may point into the rollout buffer's own rows.  n_stack in 2..8. */
int tde_step_rollout(tde_handle* h, const float* actions_dev, const uint8_t* stack_prev_dev, uint8_t* stack_next_dev,
                     int32_t n_stack, float* reward_dev, uint8_t* terminated_dev, uint8_t* truncated_dev,
                     float* info_dev, void* stream);

/* Rollout collection without moving frames ("scatter" mode): the rollout buffer holds one stacked observation per
   time slot, uint8[T + 1][E][3*n_stack][64][64] with `slot_stride_bytes` between slots.  stack_next_dev is slot t + 1.
   The frame rendered by this step is stored into every stacked observation it belongs to - the newest channel group of
   slot t + 1 and one group further down in each of the following slots_ahead - 1 slots (slots_ahead = min(n_stack,
   slots left in the buffer from t + 1 on)) - so that a slot is complete when its own step comes: 3 * n_stack * 4096
   bytes written per env and step, none read (tde_step_rollout reads (n_stack - 1) frames and writes n_stack).  For an
   env that restarted, the older groups of slot t + 1 and of the following slots are zeroed.  The caller seeds the
   older groups of the first n_stack - 1 slots of a new rollout from the observation it carries over (slot 0). */
int tde_step_rollout_scatter(tde_handle* h, const float* actions_dev, uint8_t* stack_next_dev, int64_t slot_stride_bytes,
                             int32_t slots_ahead, int32_t n_stack, float* reward_dev, uint8_t* terminated_dev,
                             uint8_t* truncated_dev, float* info_dev, void* stream);

/* VecFrameStack without moving a frame, for a stepping loop (TorchDriveVecEnv): `ring_dev` is uint8[ring_slots][E][3*n_stack][64][64],
   n_stack <= ring_slots <= 64.  The step's stacked observation is slot `ring_pos`; the caller advances ring_pos by one
   (modulo ring_slots) per step.  The new frame is stored into the newest channel group of slot ring_pos and, one group
   further down each, into the following n_stack - 1 slots of the ring, so every slot is complete when its step comes
   (tde_step_rollout_scatter on a ring).  The slot of step t stays intact for ring_slots - n_stack further steps: with
   ring_slots = n_stack + 1 the previous observation survives the next step (what an on-policy trainer needs, which
   stores the observation it acted on after stepping: examples/rl_training.py:178-181 via collect_rollouts).  Before the
   first step: tde_render_stacked into slot 0 and its frames copied one group further down into slots 1 .. n_stack - 1. */
int tde_step_stacked_ring(tde_handle* h, const float* actions_dev, uint8_t* ring_dev, int32_t ring_slots, int32_t ring_pos, int32_t n_stack,
                          float* reward_dev, uint8_t* terminated_dev, uint8_t* truncated_dev, float* info_dev, void* stream);

/* The step with the terminal observation kept (SB3 VecEnv contract of the caller, examples/rl_training.py:159:
   SubprocVecEnv stores the last observation of a finished episode in info["terminal_observation"] before it
   resets the env; off-policy trainers use it as next_obs, PPO for the time-limit bootstrap).  obs_dev and
   terminal_obs_dev are uint8[E][3*n_stack][64][64] (n_stack = 1: plain observations).  For an env that finished
   in this step, terminal_obs_dev[e] receives the frame (stack) of its final state and obs_dev[e] the first
   frame of the new episode (older stack slots zeroed); rows of the other envs in terminal_obs_dev are left
   untouched.  Needs cfg.auto_reset.  Same results as tde_step / tde_step_stacked otherwise. */
int tde_step_terminal(tde_handle* h, const float* actions_dev, uint8_t* obs_dev, int32_t n_stack, uint8_t* terminal_obs_dev,
                      float* reward_dev, uint8_t* terminated_dev, uint8_t* truncated_dev, float* info_dev, void* stream);

int tde_kinematics(tde_handle* h, const float* actions_dev, void* stream);
int tde_render(tde_handle* h, uint8_t* obs_dev, void* stream);
/* The birdview as class indices, before the palette (reference: the per-class masks that
   torchdrivesim's renderer composites into RGB; torchdriveenv/gym_env.py:123 consumes only the composite).
   nibbles_dev is uint8[E][64][32]: 4 bits per pixel, pixel (row, 2 b) in the low and (row, 2 b + 1) in the high nibble
   of byte b of the row; tde_render's RGB planes are palette[class] of exactly these classes.  It is what
   tde_step_host sends over PCIe unless cfg.host_obs_rgb = 1. */
int tde_render_classes(tde_handle* h, uint8_t* nibbles_dev, void* stream);
int tde_compute_infractions(tde_handle* h, void* stream);

/* state float[E][A][4] = x y psi v; attr float[E][A][4] = length width lr present */
int tde_get_state(tde_handle* h, float* out_dev, void* stream);
int tde_set_state(tde_handle* h, const float* in_dev, void* stream);
int tde_get_attributes(tde_handle* h, float* out_dev, void* stream);
int tde_set_attributes(tde_handle* h, const float* in_dev, void* stream);
/* float[E][A][4], columns TDE_INFR_* */
int tde_get_infractions(tde_handle* h, float* out_dev, void* stream);
/* int32[E][8]: scenario, step, target_idx, reached, light_phase, episode, map, reserved */
int tde_get_env_vars(tde_handle* h, int32_t* out_dev, void* stream);
int tde_set_env_vars(tde_handle* h, const int32_t* in_dev, void* stream);

/* Stateless kernels on caller-provided boxes (config C4).
   state_dev float[E][A][4], attr_dev float[E][A][4]; out float[E][A]. */
int tde_collision_boxes(const float* state_dev, const float* attr_dev, int32_t num_envs, int32_t num_agents,
                        float* out_count_dev, void* stream);
int tde_offroad_boxes(tde_handle* h, int32_t map_id, const float* state_dev, const float* attr_dev,
                      int32_t num_envs, int32_t num_agents, float* out_offroad_dev, void* stream);

/* Recording view: BirdviewRecordingWrapper(simulator, res=Resolution(video_res, video_res), fov=video_fov)
   gym_env.py:52-53,295-297 + simulator.get_birdviews() :174.  Draws ONE env as a width x height RGB frame
   (out_dev uint8[3][height][width], planar) from a free camera: centre (cam_x, cam_y) in world metres, heading
   cam_psi (the camera's +x axis points right in the image), `fov` metres across the width.  Same primitives,
   painter's levels, palette and fill rule as the observation; coordinates are kept in a wider fixed-point
   range so any resolution up to 4096 works.  Not part of the per-step path (two small launches per frame). */
int tde_render_view(tde_handle* h, int32_t env, float cam_x, float cam_y, float cam_psi, float fov,
                    int32_t width, int32_t height, uint8_t* out_dev, void* stream);

/* simulator.copy() :110: a second, independent handle on the same GPU holding a copy of every env (state, attributes,
   cached infractions, env variables, episode returns, frame-stack restart flags, scenario ranges, seed), made with
   device-to-device copies on `stream`.  The scenario tables are shared between the two handles (read-only,
   reference-counted: they are freed when the last handle is destroyed or uploads a new set).  Episode statistics
   start at zero in the copy. */
int tde_clone(tde_handle* h, tde_handle** out, void* stream);

/* double[TDE_NUM_STATS]; synchronises `stream`. reset_after != 0 zeroes the accumulators. */
int tde_get_episode_stats(tde_handle* h, double* out_host, int32_t reset_after, void* stream);

/* introspection used by the tests / bench */
int tde_num_kernel_launches(const tde_handle* h, int64_t* out);
int tde_device_sm_count(const tde_handle* h, int32_t* out);
/* int32[8]: road triangles, marking triangles, stop lines, grid nx, grid ny, grid list entries,
   cells flagged "every point within the offroad threshold", static render primitives after merging
   triangle pairs into convex quads */
int tde_get_map_info(const tde_handle* h, int32_t map_id, int32_t* out8);

#ifdef __cplusplus
}
#endif
#endif /* TDE_B200_H */
