/* urgym_b200.h -- C ABI of the B200-native batched simulator for UR-gym's reach-task step path.
 *
 * Drop-in boundary.  The reference (pure Python) has no FFI of its own: its hot path sits behind the Gymnasium
 * Env protocol of `RobotTaskEnv` (UR_gym/envs/core.py:222-317) and, inside it, behind `class PyBullet`
 * (UR_gym/pyb_setup.py:15), one physics client per env.  This library replaces pyb_setup.py + the arithmetic of
 * envs/robots/UR5.py and envs/tasks/reach.py for N environments at once; the Python host layer in
 * ur-gym_b200/ binds it with ctypes (see INTEGRATION.md for the stub a reference maintainer would add).
 *
 * Conventions
 *   - plain C: opaque handle, raw pointers, sizes; no torch / C++ types in any signature.
 *   - every function returns 0 on success, a negative URGYM_E* code otherwise; urgym_last_error() gives text.
 *     The library never exits or aborts; CUDA errors are reported, not swallowed.
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).  Device entry points are
 *     stream-ordered and asynchronous: they enqueue work and return without synchronising the device.
 *   - all "device" pointers are caller-owned device memory (row-major, contiguous, 16-byte aligned for the
 *     bulk-copy fast path; unaligned pointers take a slower but equivalent path).  The library allocates only its
 *     persistent structure-of-arrays state at urgym_create().
 *   - one handle per GPU per process; a handle is not thread-safe.
 *   - there is NO CPU fallback: without a CUDA device urgym_create() fails with URGYM_ENODEVICE.
 */
#ifndef URGYM_B200_H
#define URGYM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct urgym_env urgym_env_t;

/* task ids: the four registered envs in scope (UR_gym/__init__.py:19-42, UR_gym/envs/ur_tasks.py:37-90) */
enum { URGYM_TASK_ORI = 0,   /* UR5OriReach-v1  ReachOri  reach.py:141-236  obs 18, goal 6 */
       URGYM_TASK_OBS = 1,   /* UR5ObsReach-v1  ReachObs  reach.py:239-374  obs 26, goal 3 */
       URGYM_TASK_STA = 2,   /* UR5StaReach-v1  ReachSta  reach.py:377-573  obs 29, goal 6 */
       URGYM_TASK_DYN = 3 }; /* UR5DynReach-v1  ReachDyn  reach.py:576-785  obs 35, goal 6 */

/* link geometry used by getClosestPoints stand-ins (pyb_setup.py:382-456) */
enum { URGYM_GEOM_HULL = 0,     /* the reference's convex-hull links + cylinder/box scene, GJK distances */
       URGYM_GEOM_CAPSULE = 1 };/* one segment per link and for the obstacle, closed-form distances minus per-pair margins
                                   CALIBRATED against the hull geometry (ur-gym_b200/csrc/urgym_capsule_fit.h) -- an
                                   approximation, not a bound: on the hull path's own steady state the two collision
                                   flags disagree in both directions (rates: profiles/disagreement_r02.json) */

/* what `link_dist` (observation columns, reward term) measures: PyBullet.get_link_distances, pyb_setup.py:439-456 */
enum { URGYM_LD_OBSTACLE = 0,   /* links 2..6 vs the obstacle: the code the reference ships (default)                  */
       URGYM_LD_WORKBENCH = 1 };/* per link min(obstacle, table, track): what the method's docstring describes ("distance
                                   between workbench, obstacle and UR5"); the definition under which the policies shipped
                                   in Trained_Models/Trained_{Obs,Sta} (2023-09, older than the shipped code) reproduce
                                   their published success rates -- DESIGN.md section 2 */

/* state fields for urgym_get_state / urgym_set_state: the injection hooks of the reference
 * (robot.set_joint_angles core.py:161-167; task.set_goal reach.py:202-204; task.set_goal_and_obstacle
 * reach.py:328-335,483-503,702-713) and a checkpoint of the whole simulator. */
enum { URGYM_F_Q = 0,            /* float [N,6]  joint angles                                  */
       URGYM_F_GOAL = 1,         /* float [N,G]  goal (pos[,euler]); G = urgym_goal_dim()      */
       URGYM_F_OBSTACLE = 2,     /* float [N,6]  obstacle pose pos+euler (Dyn: obstacle_start) */
       URGYM_F_OBSTACLE_END = 3, /* float [N,6]  obstacle_end: Dyn (reach.py:670), Sta (18-value injection, reach.py:491-499;
                                    all zero = obstacle at rest, core.py:307-308)                     */
       URGYM_F_LINK_DIST = 4,    /* float [N,5]  link_dist == last_dist (reach.py:324,479,681) */
       URGYM_F_ELAPSED = 5,      /* int32 [N]    TimeLimit counter == ReachDyn.step_num        */
       URGYM_F_EP_RETURN = 6,    /* float [N]    running episode return                        */
       URGYM_F_VELOCITY = 7,     /* float [N,6]  ReachDyn.velocity carried over the last reset (reach.py:664-683 never clears it) */
       URGYM_F_HOT = 8,          /* float [N,24] the step kernel's copy of the episode constants + episode cache (derived
                                    from GOAL / OBSTACLE / OBSTACLE_END whenever those are set; raw words, for
                                    checkpoints: restoring it after the other fields makes a restore bit-exact) */
       URGYM_F_OBSTACLE_START = 9, /* float [N,6] Sta: ReachSta.obstacle_start (reach.py:493; the twist of the moving obstacle
                                    is (end - start) / 1 s, reach.py:518-541); Dyn: alias of URGYM_F_OBSTACLE  */
       URGYM_F_COUNT = 10 };

#define URGYM_OK 0
#define URGYM_EINVAL (-1)     /* bad argument                         */
#define URGYM_ENODEVICE (-2)  /* no usable CUDA device                */
#define URGYM_ECUDA (-3)      /* a CUDA runtime call failed           */
#define URGYM_ENOMEM (-4)     /* allocation failed                    */
#define URGYM_EUNSUPPORTED (-5)

#define URGYM_MAX_EPISODE_STEPS 100   /* register(max_episode_steps=100)  UR_gym/__init__.py:22,28,34,41 */
#define URGYM_STATS_COUNT 8

/* ---- lifetime ------------------------------------------------------------------------------------------- */
/* n_envs environments with global indices [env_index_offset, env_index_offset + n_envs) on CUDA device `device`.
 * `seed` keys the counter-based reset stream (Philox4x32-10): the draws of a reset depend on (seed, global env
 * index, reset event), where the reset event is the handle's count of urgym_step/urgym_reset calls so far -- not on
 * the sharding, so 1/2/4/8-GPU runs that issue the same calls produce the same episodes.
 * n_envs < 2^31 per handle (URGYM_EINVAL otherwise; env_index_offset is 64-bit: shard larger batches over handles).
 * All envs start un-reset: call urgym_reset(h, NULL, ...) once.   Replaces: gymnasium.make(id)
 * (ur_tasks.py:37-90: PyBullet(...) + UR5Ori(...) + Reach*(...)). */
int urgym_create(urgym_env_t **out, int task, int geom, int64_t n_envs, int64_t env_index_offset,
                 uint64_t seed, int device);
int urgym_destroy(urgym_env_t *h);                       /* RobotTaskEnv.close  core.py:319-320 */
const char *urgym_last_error(const urgym_env_t *h);      /* h may be NULL: error of the last failed create */

int urgym_obs_dim(int task);    /* 18 / 26 / 29 / 35 */
int urgym_goal_dim(int task);   /* 6 / 3 / 6 / 6     */
int64_t urgym_num_envs(const urgym_env_t *h);

/* ---- the hot path ---------------------------------------------------------------------------------------- */
/* One env step for every env, then auto-reset of the envs that finished (SB3 DummyVecEnv semantics, which is what
 * train.py's loop sees: train.py:39-60).            Replaces: RobotTaskEnv.step  core.py:303-317 (+ TimeLimit)
 *   actions      [N,6] float  in   clipped to [-1,1] inside                              UR5.py:273-279
 *   obs          [N,D] float  out  next observation; for finished envs the first observation of the new episode
 *   achieved     [N,G] float  out  achieved_goal belonging to `obs`
 *   desired      [N,G] float  out  desired_goal belonging to `obs` (may be NULL: read URGYM_F_GOAL instead)
 *   reward       [N]   float  out
 *   terminated   [N]   uint8  out  success or collision                                  core.py:313
 *   truncated    [N]   uint8  out  elapsed >= 100 (TimeLimit)                            UR_gym/__init__.py
 *   is_success   [N]   uint8  out  info["is_success"]                                    core.py:315
 *   terminal_obs [N,D] float  out  rows of finished envs only: the observation of the final step
 *                                  (DummyVecEnv's info["terminal_observation"]); may be NULL
 *   terminal_achieved [N,G] float out  likewise; may be NULL                                                    */
int urgym_step(urgym_env_t *h, const float *actions, float *obs, float *achieved, float *desired, float *reward,
               uint8_t *terminated, uint8_t *truncated, uint8_t *is_success, float *terminal_obs,
               float *terminal_achieved, void *stream);

/* The same step for the envs [first, first + count) only (array pointers are those of the WHOLE arrays).  Lets a caller
 * advance disjoint env ranges as independent "chains" on different streams, so that one range's auto-reset kernel
 * overlaps another range's step kernel.  Each chain (0..7) counts its own reset events: a chain must be stepped as
 * often as the others before the next whole-batch urgym_step / urgym_reset, and always with the same range.
 * Results are identical to urgym_step over the whole batch. */
int urgym_step_range(urgym_env_t *h, int64_t first, int64_t count, int chain, const float *actions, float *obs,
                     float *achieved, float *desired, float *reward, uint8_t *terminated, uint8_t *truncated,
                     uint8_t *is_success, float *terminal_obs, float *terminal_achieved, void *stream);

/* Reset the envs whose mask byte is non-zero (mask == NULL: all).  Writes the first observation of the new episode
 * into the rows of obs/achieved/desired it resets (any of them may be NULL).
 *                                                      Replaces: RobotTaskEnv.reset  core.py:263-273 */
int urgym_reset(urgym_env_t *h, const uint8_t *mask, float *obs, float *achieved, float *desired, void *stream);

/* Recompute obs / achieved / desired from the current state without stepping (any may be NULL).
 *                                                      Replaces: RobotTaskEnv._get_obs  core.py:252-261 */
int urgym_observe(urgym_env_t *h, float *obs, float *achieved, float *desired, void *stream);

/* After injecting goal / obstacle / joints with urgym_set_state: recompute link_dist (= last_dist) and report the
 * collision flag, as the tail of set_goal_and_obstacle does (reach.py:333-335,501-503,711-713).
 * collision [N] uint8 out, may be NULL. */
int urgym_refresh(urgym_env_t *h, uint8_t *collision, void *stream);

/* ---- state access (device pointers, stream-ordered) ------------------------------------------------------- */
int urgym_get_state(urgym_env_t *h, int field, void *dst, void *stream);
int urgym_set_state(urgym_env_t *h, int field, const void *src, void *stream);

/* Episode statistics accumulated on the device since the last call with reset_after != 0; synchronises `stream`.
 *   out[0] episodes finished   out[1] sum of episode returns   out[2] sum of episode lengths
 *   out[3] successes           out[4] collisions               out[5] truncations
 *   out[6] env steps taken     out[7] reset rejection iterations
 * These are the per-shard sums that the host layer all-reduces over ranks. */
int urgym_stats(urgym_env_t *h, double out[URGYM_STATS_COUNT], int reset_after, void *stream);

/* ---- host-buffer entry point (what a CPU-side caller such as SB3's DummyVecEnv replacement uses) ----------- */
/* Same as urgym_step but every pointer is HOST memory (pageable or pinned); copies actions to the device, steps,
 * copies the results back and synchronises.  desired/terminal_* may be NULL. */
int urgym_step_host(urgym_env_t *h, const float *actions, float *obs, float *achieved, float *desired,
                    float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *is_success,
                    float *terminal_obs, float *terminal_achieved);
int urgym_reset_host(urgym_env_t *h, const uint8_t *mask, float *obs, float *achieved, float *desired);
/* The same step without the final wait, for callers that keep two sets of (pinned) host buffers: step k goes into
 * `slot` k % 2 and the call returns as soon as everything is enqueued; urgym_host_wait(h, slot) blocks until that slot's
 * results are complete in host memory.  Enqueueing step k + 1 (other slot) before waiting for step k lets its kernels run
 * under step k's device-to-host copies.  A slot must be waited for before it is used again (URGYM_EINVAL otherwise), and
 * its `actions` buffer must stay untouched until then.  Do not mix with urgym_step_host while a slot is in flight. */
int urgym_step_host_async(urgym_env_t *h, int slot, const float *actions, float *obs, float *achieved, float *desired,
                          float *reward, uint8_t *terminated, uint8_t *truncated, uint8_t *is_success,
                          float *terminal_obs, float *terminal_achieved);
int urgym_host_wait(urgym_env_t *h, int slot);

/* ---- device-resident replay ring (the env side of train.py's SAC loop without leaving the GPU) --------------------- */
/* Append one transition per env to caller-owned device rings of `capacity` rows (capacity >= N):
 *   ring_obs / ring_next_obs [capacity, D], ring_actions [capacity, 6], ring_reward [capacity], ring_done,
 *   ring_timeout [capacity] uint8.   Row (cursor + i) % capacity receives env i's (obs, action, reward, next_obs, done,
 *   TimeLimit.truncated); next_obs is terminal_obs[i] where the env finished (SB3 VecEnv / ReplayBuffer semantics:
 *   train.py:39-48 through DummyVecEnv).  `cursor` is a device-resident counter that the call advances by N, so a captured
 *   CUDA graph of (policy, urgym_step, urgym_replay_write) keeps appending on every replay.
 *   obs = the observation the actions were computed from; next_obs / terminal_obs / reward / flags = urgym_step's outputs. */
int urgym_replay_write(urgym_env_t *h, const float *obs, const float *actions, const float *reward,
                       const uint8_t *terminated, const uint8_t *truncated, const float *next_obs,
                       const float *terminal_obs, float *ring_obs, float *ring_next_obs, float *ring_actions,
                       float *ring_reward, uint8_t *ring_done, uint8_t *ring_timeout, int64_t capacity,
                       unsigned long long *cursor, void *stream);

/* ---- options / checkpoint scalars ---------------------------------------------------------------------------- */
/* auto-reset on (default, DummyVecEnv semantics) or off (a bare RobotTaskEnv: finished envs keep their state and
 * keep stepping until the caller resets them, core.py:303-317). */
int urgym_set_autoreset(urgym_env_t *h, int enabled);
/* URGYM_LD_*; takes effect with the next launch: call urgym_refresh afterwards so that link_dist = last_dist are
 * re-measured in the new mode.  URGYM_EUNSUPPORTED for UR5OriReach (no obstacle, no link_dist). */
int urgym_set_link_dist_mode(urgym_env_t *h, int mode);
/* the reset-event counter (position of the reset stream; the largest of the chains' counters); together with the
 * URGYM_F_* fields it checkpoints a handle.  urgym_set_event sets every chain's counter. */
int urgym_get_event(const urgym_env_t *h, uint32_t *event);
/* Raise every chain's reset-event counter to the largest of them (stream-ordered, one tiny kernel).  To be enqueued before
 * a group of urgym_step_range calls whenever the chains may have been stepped unequally before (a different chain layout,
 * a graph with fewer chains): afterwards no chain can draw from a reset-stream position that was already used. */
int urgym_sync_events(urgym_env_t *h, void *stream);
/* re-key the reset stream: RobotTaskEnv.reset(seed=...) re-creates task.np_random (core.py:267) */
int urgym_set_seed(urgym_env_t *h, uint64_t seed);
int urgym_set_event(urgym_env_t *h, uint32_t event);

/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t urgym_launch_count(const urgym_env_t *h);

/* Kernel timing for the roofline report.  While enabled, urgym_step / urgym_step_range record CUDA events on the
 * launching stream directly before and after the step kernel and after the auto-reset kernel (up to 256 steps are
 * kept; not to be used while a stream is being captured).  urgym_profile_read synchronises the device and returns
 * the average duration of the two kernels in milliseconds over the recorded steps, then clears the record. */
int urgym_profile_enable(urgym_env_t *h, int enabled);
int urgym_profile_read(urgym_env_t *h, double *step_kernel_ms, double *reset_kernel_ms, int *steps);

/* ---- the motor-driven robot path ------------------------------------------------------------ */
/* `UR5IAIReach-v1` (UR_gym/envs/ur_tasks.py:10-21): robot `UR5` (UR_gym/envs/robots/UR5.py:10-118, urdf/ur5.urdf), task
 * `ReachIAI` (UR_gym/envs/tasks/reach.py:9-66), N envs at once.  The one registered env of the reference whose step goes
 * through setJointMotorControlArray(POSITION_CONTROL) (pyb_setup.py:365-380) and 20 dynamic substeps (pyb_setup.py:52-55)
 * instead of a joint teleport.  Observation [N,6] = ee position + ee linear velocity (UR5.py:92-97), goals [N,3].
 * Bullet's substep (forward dynamics, motor rows, 50 Gauss-Seidel sweeps, semi-implicit Euler) is restated from its
 * published algorithm and is NOT pinned to a PyBullet run: oracle/ur_motor_oracle.c lists what is recalled; joint-limit
 * and contact rows are not modelled.  A handle of its own: nothing here touches urgym_env_t. */
typedef struct urgym_motor urgym_motor_t;
enum { URGYM_MOTOR_F_Q = 0,        /* float [N,6] joint angles      (robot.set_joint_angles, core.py:161-167) */
       URGYM_MOTOR_F_QD = 1,       /* float [N,6] joint velocities                                            */
       URGYM_MOTOR_F_GOAL = 2,     /* float [N,3] goal              (task.set_goal)                           */
       URGYM_MOTOR_F_ELAPSED = 3,  /* int32 [N]   TimeLimit counter                                           */
       URGYM_MOTOR_F_COUNT = 4 };
#define URGYM_MOTOR_OBS_DIM 6
#define URGYM_MOTOR_GOAL_DIM 3
int urgym_motor_create(urgym_motor_t **out, int64_t n_envs, int64_t env_index_offset, uint64_t seed, int device);
int urgym_motor_destroy(urgym_motor_t *h);
const char *urgym_motor_last_error(const urgym_motor_t *h);
/* RobotTaskEnv.reset (core.py:263-273): neutral pose at rest (UR5.py:99-103), new goal (reach.py:48-57); mask NULL = all */
int urgym_motor_reset(urgym_motor_t *h, const uint8_t *mask, float *obs, float *achieved, float *desired, void *stream);
/* RobotTaskEnv.step (core.py:303-317) + TimeLimit(100), then auto-reset of the finished envs (terminal rows kept) */
int urgym_motor_step(urgym_motor_t *h, const float *actions, float *obs, float *achieved, float *desired, float *reward,
                     uint8_t *terminated, uint8_t *truncated, uint8_t *is_success, float *terminal_obs, void *stream);
int urgym_motor_get_state(urgym_motor_t *h, int field, void *dst, void *stream);
int urgym_motor_set_state(urgym_motor_t *h, int field, const void *src, void *stream);
int urgym_motor_stats(urgym_motor_t *h, double *out8, int reset);   /* same layout as urgym_stats */
int64_t urgym_motor_launch_count(const urgym_motor_t *h);

#ifdef __cplusplus
}
#endif
#endif /* URGYM_B200_H */
