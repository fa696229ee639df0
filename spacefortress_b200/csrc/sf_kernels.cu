// sf_kernels.cu — kernels and the C-ABI (include/sf_b200.h) of the B200-native batched Space Fortress
// simulator. sm_100a only; no CPU fallback: every entry point fails with SF_ERR_CUDA without a device.
//
// Kernels
//   sf_rollout_kernel           fused step + render + auto-reset for T consecutive ticks (T=1 == sf_step). One block
//                               of 24 warps per SM owns groups of <= 32 envs: warp 0 steps them one env per lane
//                               (SoA, 128-bit loads/stores) up to two ticks ahead and prepares the next stage, the
//                               other warps run the block-cooperative frame pipeline of sf_render.cuh and stream the
//                               84x84 frames to HBM.
//                               Groups beyond the first 148 are handed out first come first served (SfSched); between
//                               the ticks of a launch the group's scalars stay in shared memory (SfStepSmem); the static
//                               tables arrive as one bulk copy of a per-handle image (sf_pack_static_kernel).
//   sf_step_only_kernel         state-only variant (render off), one env per thread.
//   sf_features_kernel          the feature observation types of ssf_env.py:95-157, one env per thread.
//   sf_policy_input_kernel      4-frame stack -> the policy's first-layer input (bf16 or fp32, space-to-depth NHWC).
//   sf_reset_kernel / sf_seed_kernel / sf_render_kernel / sf_get_state_kernel / sf_set_state_kernel / sf_pack_static_kernel
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/sf_b200.h"
#include "sf_geom.h"
#include "sf_render.cuh"
#include "sf_state.cuh"
#include "sf_step.cuh"
#include "sf_tables.h"

#define SF_WARPS_PER_BLOCK SF_RENDER_WARPS  // sf_render.cuh: one block per SM
#define SF_BLOCK (32 * SF_WARPS_PER_BLOCK)
#ifndef SF_BLOCKS_PER_SM
#define SF_BLOCKS_PER_SM 1  // resident blocks per SM (each with its own stepping warp, pools and barriers)
#endif

// ------------------------------------------------------------------------------------------------
// synthetic policy: stateless counter hash; same function on host and device
// ------------------------------------------------------------------------------------------------
__host__ __device__ inline unsigned sf_hash3(unsigned seed, unsigned long long env, unsigned long long t) {
  unsigned long long z = (env * 0x9E3779B97F4A7C15ull) ^ (t * 0xC2B2AE3D27D4EB4Full) ^ ((unsigned long long)seed << 32 | 0x5Fu);
  z ^= z >> 30; z *= 0xBF58476D1CE4E5B9ull;
  z ^= z >> 27; z *= 0x94D049BB133111EBull;
  z ^= z >> 31;
  return (unsigned)(z >> 32);
}
__host__ __device__ inline int sf_hash_action(unsigned seed, unsigned long long env, unsigned long long t, int num_actions) {
  return (int)(((unsigned long long)sf_hash3(seed, env, t) * (unsigned)num_actions) >> 32);
}

// ------------------------------------------------------------------------------------------------
// finished-episode statistics: warp-reduced, then one atomic per field per warp
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ long long sf_warp_sum_ll(long long v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Finished-episode statistics of the `finished` lanes of a warp (called by all 32 lanes). Out of line and by VALUE: a
// reference to the env's registers handed to a function that is not inlined forces the whole SfEnv into local memory in
// the caller (measured: the state-only kernel lost 30 %).
__device__ __noinline__ void sf_accumulate_episode_vals(unsigned long long* epi, bool finished, int lane, int ret_i, int length, int points_bits, int raw_bits,
                                                        int max_vlner, int4 st0, int4 st1, int4 st2) {
  long long f[SF_NUM_EPISODE_STATS];
  long long ret = ret_i;
  f[0] = 1; f[1] = ret; f[2] = ret * ret; f[3] = length;
  f[4] = st0.x; f[5] = st0.y; f[6] = st0.z; f[7] = st0.w;
  f[8] = st1.x; f[9] = st1.y; f[10] = st1.z; f[11] = st1.w;
  f[12] = st2.x; f[13] = st2.y; f[14] = st2.z; f[15] = st2.w; f[16] = max_vlner;
  f[17] = (long long)__float2int_rz(__int_as_float(points_bits));
  f[18] = __double2ll_rn((double)__int_as_float(raw_bits) * 1000.0);
  f[19] = st1.y;  // fortress kills of the episode (== destroyedFortresses; rl/train.py:81 sums info)
  f[20] = 0; f[21] = 0; f[22] = 0; f[23] = 0;
#pragma unroll
  for (int k = 0; k < 20; k++) {
    long long v = sf_warp_sum_ll(finished ? f[k] : 0);
    if (lane == 0 && v) atomicAdd(&epi[k], (unsigned long long)v);
  }
  int mv = finished ? max_vlner : 0;
  mv = sf_warp_max(mv);
  if (lane == 0 && mv) atomicMax(&epi[20], (unsigned long long)mv);
}
// the finished lanes' pending Stats increments go to the arrays first (reductions: read back through L2)
__device__ __forceinline__ void sf_accumulate_episode(const SfDev& D, int env, SfEnv& e, bool finished, int lane) {
  int4 st0 = make_int4(0, 0, 0, 0), st1 = st0, st2 = st0;
  if (finished) { sf_flush_stats(D, env, e); st0 = __ldcg(&D.st0[env]); st1 = __ldcg(&D.st1[env]); st2 = __ldcg(&D.st2[env]); }
  sf_accumulate_episode_vals(D.epi, finished, lane, e.q3.w, e.q3.z, e.q3.x, e.q3.y, e.st3.x, st0, st1, st2);
}
// The rollout kernel's stepping warp takes the env by REFERENCE instead, out of line: that keeps its SfEnv in local memory
// (L1) and its register demand low, and with it the register allocation of the whole kernel body — the drawing code
// inlined next to it then comes out without a single spill (with the by-value call it has ~15 local loads / stores in
// sf_draw_stage / sf_phase_strokes / sf_env_base_patch and the launch is 5 % slower: tools/gpu_ab.py + nvdisasm -g).
__device__ __noinline__ void sf_accumulate_episode_ref(const SfDev& D, int env, SfEnv& e, bool finished, int lane) {
  sf_accumulate_episode(D, env, e, finished, lane);
}

// the renderer's view of one stepped env (written by the env's lane into the block's record array)
__device__ __forceinline__ void sf_make_env_rec(const SfEnv& e, int env, unsigned shell_vis, SfEnvRec& r) {
  r.px = e.pos.x; r.py = e.pos.y;
  r.core = (unsigned)e.q0.x; r.pmask = (unsigned)e.q0.y;
  r.points_i = __float2int_rz(__int_as_float(e.q3.x));
  r.vuln = e.q1.z;
  r.kill_bar = (e.q1.z > 10 && e.q1.y < 250) ? 1 : 0;
  r.env = env;
  r.s0 = 0; r.ebox = 0;
  if (!(r.core & SF_CORE_SHIP_ALIVE)) {  // explosion sprite box (draw.cpp:235-237)
    const SfPt c = sf_xform_base(r.px, r.py);
    r.ebox = (((c.x >> 8) - 13) + 64) | ((((c.y >> 8) - 13) + 64) << 8);
  }
  r.life = (unsigned)e.st3.z + 1u;
  r.building = 0;  // decided by sf_publish_recs
  r.ns = sf_count_strokes(r.core, r.pmask, shell_vis);
  r.shell_vis = (int)shell_vis;
}

// Device-resident scheduling state of sf_rollout_kernel (one per handle; zero at creation, self re-arming): with more
// groups than blocks, the groups after the first gridDim.x are handed out first come first served (the groups differ
// in cost and the launch ends with the slowest block: +5..8 % at 65 536 envs over static striding).
// (Tried and dropped: dealing the envs of a one-group-per-block launch to the blocks by a cost key read from their
// state, so that every block gets the same mix of dead and live ships — the spread of the block times fell only from
// 9.8 % to 7.7 % and the sort at the end of every launch cost more than that gained.)
struct SfSched {
  int next_group, done;                 // groups handed out beyond the first gridDim.x; blocks that have finished
};

struct SfRollArgs {
  int T, EB, ngroups, flags;  // EB = envs per group (<= 32), ngroups = ceil(n / EB)
  const int* actions;  // [T][n] or NULL
  unsigned action_seed;
  long long t0;
  unsigned char* obs;  // [T][n][obs_bytes]
  int* reward;         // [T][n]
  unsigned char* done;
  unsigned char* fortkill;
  unsigned* events;
  SfSched* sched;      // NULL: static assignment (sf_render_kernel-like contiguous groups, no hand-out)
  int env0, envn;      // the envs this launch steps: [env0, env0 + envn) (sf_step_host steps the slab in slices)
  // SF_FLAG_HOST_DELTA (sf_step_host, T == 1): every block ends by sending the changed granules of its own frames
  uint4* host_obs;     // device alias of the caller's page-locked observation buffer, or NULL
  uint4* mirror;       // device copy of what that buffer holds
  unsigned long long* delta_stats;
  int delta_lanes;     // granule in 16-byte lanes
};

// ------------------------------------------------------------------------------------------------
// fused step + render
// ------------------------------------------------------------------------------------------------
// One block renders groups of EB envs (persistent over the groups it owns, group-major, T ticks each). Per tick:
// warp 0 steps the group (one env per lane: SoA 128-bit loads/stores) and publishes the env records; then all
// warps run the block-cooperative frame pipeline (sf_render.cuh).
// one tick of a group (warp 0, one env per lane): step, outputs, auto-reset, staged env record
__device__ __noinline__ void sf_step_group(const SfDev& D, const SfRollArgs& A, int group, int t, SfEnvRec* recs) {
  const int lane = threadIdx.x & 31;
  const SfHot* H = &sf_block_smem().hot;  // the step's tables from shared memory
  const bool autoreset = !(A.flags & SF_FLAG_NO_AUTORESET);
  const int env = A.env0 + group * A.EB + lane;
  const bool mine = lane < A.EB && env < A.env0 + A.envn;
  SfEnv e;
  bool finished = false;
  unsigned shell_vis = 0;
#ifdef SF_BARRIER_TIMING
  unsigned stmask_ = __ballot_sync(0xffffffffu, mine);
#endif
  SF_ST_BEGIN();
  SfStepSmem& SS = sf_block_smem().step;
  const bool first_tick = t == 0, last_tick = t == A.T - 1;
  if (mine) {
    if (first_tick) sf_load_env(D, env, e);
    else {  // between the ticks of a launch the group's scalars live in shared memory
      e.pos = SS.pos[lane]; e.vel = SS.vel[lane];
      e.q0 = SS.q0[lane]; e.q1 = SS.q1[lane]; e.q2 = SS.q2[lane]; e.q3 = SS.q3[lane]; e.st3 = SS.st3[lane];
      e.d0 = SS.d0[lane]; e.d1 = SS.d1[lane]; e.d2 = SS.d2[lane];
    }
#ifdef SF_BARRIER_TIMING
    if (e.q0.x == 0x7fffffff) e.q0.y = 0;  // (a use of the loaded words: the section ends when they have arrived)
#endif
    SF_ST(8);
    int a = A.actions ? A.actions[(size_t)t * D.n + env]
                      : sf_hash_action(A.action_seed, (unsigned long long)(D.first_global_env + env), (unsigned long long)(A.t0 + t), D.num_actions);
    int km = (A.flags & SF_FLAG_ACTIONS_ARE_KEYMASKS) ? (a & 15) : D.keymask_of_action[min(max(a, 0), D.num_actions - 1)];
    SfStepOut o;
#ifdef SF_BARRIER_TIMING
    o.sync_mask = stmask_;
#endif
    sf_env_step(D, H, env, e, km, autoreset, (A.flags & SF_FLAG_RAW_REWARD) != 0, o);
    size_t oi = (size_t)t * D.n + env;
    if (A.reward) A.reward[oi] = o.reward;
    if (A.done) A.done[oi] = o.done;
    if (A.fortkill) A.fortkill[oi] = o.fort_kill;
    if (A.events) A.events[oi] = o.events;
    finished = o.done && autoreset;
    shell_vis = o.shell_vis;
  }
#ifdef SF_BARRIER_TIMING
  ts_ = clock64(); stmask_ = 0xffffffffu;
#endif
  if (__any_sync(0xffffffffu, finished)) {
#ifdef SF_STEP_ENV_BY_REF
    sf_accumulate_episode_ref(D, env, e, finished, lane);
#else
    sf_accumulate_episode(D, env, e, finished, lane);
#endif
    if (finished) { sf_new_game(D, H, env, e); shell_vis = 0; }  // gym_vecenv: the returned obs is the first frame of the new episode
  }
  if (mine) {
    if (last_tick) sf_store_env(D, env, e);
    else {
      SS.pos[lane] = e.pos; SS.vel[lane] = e.vel;
      SS.q0[lane] = e.q0; SS.q1[lane] = e.q1; SS.q2[lane] = e.q2; SS.q3[lane] = e.q3; SS.st3[lane] = e.st3;
      if ((t & 7) == 7) sf_flush_stats(D, env, e);  // the pending Stats increments are 8-bit fields: every 8 ticks is often enough (as in sf_step_only_kernel)
      SS.d0[lane] = e.d0; SS.d1[lane] = e.d1; SS.d2[lane] = e.d2;
    }
    sf_make_env_rec(e, env, shell_vis, recs[lane]);
  } else recs[lane].env = -1;
  __syncwarp();
  SF_ST(15);
}


// The two roles of the rollout kernel, out of line (separate register allocations, see sf_stepper_ticks). Both walk the
// same sequence of groups: the block's first group is blockIdx.x, further ones are handed out first come first served
// (the groups differ in cost, and so do the SMs' speeds). Warp 0 fetches the id of the NEXT group when it starts a group
// and leaves it in shared memory (two slots, alternating); the other warps read it when they are done with the current
// group, many barriers later. The step of tick t + 1 (warp 0) runs while the other warps draw tick t, across groups too.
__device__ __noinline__ void sf_rollout_stepper(const SfDev& D, const SfRollArgs& A) {
  const int lane = threadIdx.x & 31;
  SfBlockSmem& B = sf_block_smem();
  const bool native = (A.flags & SF_FLAG_NATIVE_OBS) != 0;
  int stage = 0, group = blockIdx.x;
#pragma unroll 1
  for (int k = 1; group < A.ngroups; k ^= 1) {
    if (lane == 0) B.next_group[k] = A.sched ? (int)gridDim.x + atomicAdd(&A.sched->next_group, 1) : group + (int)gridDim.x;
    sf_stepper_ticks(D, B, lane, native, A.T, stage, [&](int t, SfTeamSmem& Tm, int h) { sf_step_group(D, A, group, t, &Tm.env[32 * h]); });
    group = B.next_group[k];
  }
  // the last block to leave re-arms the counters for the next launch (every block has made its last fetch by then)
  if (lane == 0 && A.sched) {
    __threadfence();
    if (atomicAdd(&A.sched->done, 1) == (int)gridDim.x - 1) { A.sched->next_group = 0; A.sched->done = 0; __threadfence(); }
  }
}
__device__ __forceinline__ void sf_rollout_drawer(const SfDev& D, const SfRollArgs& A) {
  const int lane = threadIdx.x & 31;
  SfBlockSmem& B = sf_block_smem();
  SfWarpSmem& W = sf_my_smem();
  SfFrameOut out;
  out.native = (A.flags & SF_FLAG_NATIVE_OBS) ? 1 : 0;
  out.obs_bytes = out.native ? (size_t)SF_NAT_H * SF_NAT_W : (size_t)84 * 84;
  out.obs = A.obs;
  out.tick_bytes = (size_t)D.n * out.obs_bytes;
  SfStageState st;
  st.stage = 0; st.prev_used = 0;
  int group = blockIdx.x;
#pragma unroll 1
  for (int k = 1; group < A.ngroups; k ^= 1) {
    sf_drawer_ticks(D, B, W, lane, out, A.T, st);
    group = B.next_group[k];
  }
}

// SF_FLAG_HOST_DELTA inside the kernel: the block compares the frames of its own groups (just written: L2) with the
// mirror, 16 bytes per thread and four in flight, and stores the granules that differ straight into the host buffer
// (posted writes across PCIe) and into the mirror. Blocks finish at different times (at T = 1 the median block is done
// after ~70 % of the launch), so most of this and most of the PCIe traffic hides under the slower blocks; a separate
// kernel (sf_host_delta_kernel, the fallback for frames that are not a multiple of 16 bytes) would start after the
// slowest block.
__device__ __noinline__ void sf_block_host_delta(const SfDev& D, const SfRollArgs& A) {
  const unsigned lane = threadIdx.x & 31;
  const unsigned gmask = (1u << A.delta_lanes) - 1u, gsh = lane & ~(unsigned)(A.delta_lanes - 1);
  const uint4* cur = reinterpret_cast<const uint4*>(A.obs);
  const size_t per16 = (size_t)(84 * 84) / 16;
  unsigned sent = 0;
#pragma unroll 1
  for (int group = blockIdx.x; group < A.ngroups; group += gridDim.x) {
    const int env_lo = A.env0 + group * A.EB, env_hi = min(env_lo + A.EB, A.env0 + A.envn);
    const size_t c1 = (size_t)env_hi * per16;
#pragma unroll 1
    for (size_t base = (size_t)env_lo * per16; base < c1; base += 4 * SF_BLOCK) {
      uint4 a[4], b[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t i = base + u * SF_BLOCK + threadIdx.x;
        a[u] = b[u] = make_uint4(0, 0, 0, 0);
        if (i < c1) { a[u] = __ldcg(cur + i); b[u] = __ldcg(A.mirror + i); }
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const size_t i = base + u * SF_BLOCK + threadIdx.x;
        const bool diff = (a[u].x != b[u].x) | (a[u].y != b[u].y) | (a[u].z != b[u].z) | (a[u].w != b[u].w);
        const unsigned m = __ballot_sync(0xffffffffu, diff);
        if (m == 0) continue;
        if (i < c1 && ((m >> gsh) & gmask)) { A.host_obs[i] = a[u]; sent++; }
        if (diff) A.mirror[i] = a[u];
      }
    }
  }
  for (int o = 16; o; o >>= 1) sent += __shfl_xor_sync(0xffffffffu, sent, o);
  if (lane == 0 && sent) atomicAdd(&A.delta_stats[0], (unsigned long long)sent * 16u);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(&A.delta_stats[1], 1ull);
}

// (HOST_DELTA is a second instantiation, used by sf_step_host only: the register allocation of the plain kernel is
// sensitive to anything that is added to its body)
template <bool HOST_DELTA>
__global__ void __launch_bounds__(SF_BLOCK, SF_BLOCKS_PER_SM) sf_rollout_kernel(const __grid_constant__ SfDev D, const __grid_constant__ SfRollArgs A) {
  sf_block_smem_init(D.static_image);
  sf_warp_smem_init(sf_my_smem(), threadIdx.x & 31);
  if (threadIdx.x < 32) sf_rollout_stepper(D, A); else sf_rollout_drawer(D, A);
  if (HOST_DELTA) {
    __syncthreads();  // every frame of this block's groups is written (the bulk copies were waited for before the patches)
    sf_block_host_delta(D, A);
  }
}

// state-only: one env per thread
__global__ void __launch_bounds__(128) sf_step_only_kernel(SfDev D, SfRollArgs A) {
  const int env = A.env0 + blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool mine = env < A.env0 + A.envn;
  const bool autoreset = !(A.flags & SF_FLAG_NO_AUTORESET);
  SfEnv e;
  if (mine) sf_load_env(D, env, e);
  for (int t = 0; t < A.T; t++) {
    bool finished = false;
    if (mine) {
      int a = A.actions ? A.actions[(size_t)t * D.n + env]
                        : sf_hash_action(A.action_seed, (unsigned long long)(D.first_global_env + env), (unsigned long long)(A.t0 + t), D.num_actions);
      int km = (A.flags & SF_FLAG_ACTIONS_ARE_KEYMASKS) ? (a & 15) : D.keymask_of_action[min(max(a, 0), D.num_actions - 1)];
      SfStepOut o;
#ifdef SF_BARRIER_TIMING
      o.sync_mask = __activemask();
#endif
      sf_env_step(D, &D.tab->hot, env, e, km, autoreset, (A.flags & SF_FLAG_RAW_REWARD) != 0, o);
      size_t oi = (size_t)t * D.n + env;
      if (A.reward) A.reward[oi] = o.reward;
      if (A.done) A.done[oi] = o.done;
      if (A.fortkill) A.fortkill[oi] = o.fort_kill;
      if (A.events) A.events[oi] = o.events;
      finished = o.done && autoreset;
    }
    if (__any_sync(0xffffffffu, finished)) {
      sf_accumulate_episode(D, env, e, finished, lane);
      if (finished) sf_new_game(D, &D.tab->hot, env, e);
    }
    if (mine && (t & 7) == 7) sf_flush_stats(D, env, e);  // the pending increments are 8-bit fields
  }
  if (mine) sf_store_env(D, env, e);
}

// render the current state (Game.draw): the same block-cooperative pipeline without the step
__global__ void __launch_bounds__(SF_BLOCK, SF_BLOCKS_PER_SM) sf_render_kernel(SfDev D, unsigned char* obs, int flags, const unsigned char* mask, int EB, int ngroups) {
  const int lane = threadIdx.x & 31;
  SfBlockSmem& B = sf_block_smem();
  SfWarpSmem& W = sf_my_smem();
  sf_block_smem_init(D.static_image);
  sf_warp_smem_init(W, lane);
  const bool native = (flags & SF_FLAG_NATIVE_OBS) != 0;
  if (threadIdx.x < 32) {
    int stage = 0;
#pragma unroll 1
    for (int group = blockIdx.x; group < ngroups; group += gridDim.x)
      sf_stepper_ticks(D, B, lane, native, 1, stage, [&](int, SfTeamSmem& Tm, int) {
        const int env = group * EB + lane;
        const bool mine = lane < EB && env < D.n && (!mask || mask[env]);
        if (mine) {
          SfEnv e;
          sf_load_env(D, env, e);
          sf_make_env_rec(e, env, sf_visible_shells(D, env, (unsigned)e.q0.y), Tm.env[lane]);
        } else Tm.env[lane].env = -1;
        __syncwarp();
      });
  } else {
    SfFrameOut out;
    out.native = native ? 1 : 0;
    out.obs_bytes = out.native ? (size_t)SF_NAT_H * SF_NAT_W : (size_t)84 * 84;
    out.obs = obs;
    out.tick_bytes = 0;
    SfStageState st;
    st.stage = 0; st.prev_used = 0;
#pragma unroll 1
    for (int group = blockIdx.x; group < ngroups; group += gridDim.x) sf_drawer_ticks(D, B, W, lane, out, 1, st);
  }
}

// once per handle (and after sf_set_glyph_masks): the static part of SfBlockSmem, built from the tables, as an image in
// global memory that the rendering kernels fetch with one bulk copy
__global__ void __launch_bounds__(SF_BLOCK) sf_pack_static_kernel(SfDev D, unsigned char* image) {
  sf_block_smem_fill(D.tab);
  __syncthreads();
  const int4* src = reinterpret_cast<const int4*>(&sf_block_smem());
  for (int k = threadIdx.x; k < (int)(SF_STATIC_BYTES / 16); k += blockDim.x) reinterpret_cast<int4*>(image)[k] = src[k];
}

__global__ void sf_seed_kernel(SfDev D, const unsigned* seeds) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= D.n) return;
  SfEnv e;
  sf_load_env(D, env, e);
  sf_srand(D, env, e, seeds ? seeds[env] : 1u);
  D.expstamp[env] = 0;  // the rand() call counter restarts: no cached explosion belongs to a life of the new stream
  D.expo_meta[env] = make_uint2(0u, 0u);
  sf_store_env(D, env, e);
}

__global__ void sf_set_ticks_kernel(SfDev D, const int* ticks) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= D.n) return;
  int4 q = D.q3[env];
  q.z = max(ticks[env], 0);
  D.q3[env] = q;
}

__global__ void sf_reset_kernel(SfDev D, const unsigned char* mask, int clear_prev_vlner) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= D.n) return;
  if (mask && !mask[env]) return;
  SfEnv e;
  sf_load_env(D, env, e);
  if (clear_prev_vlner) e.q1.w = 0;
  sf_new_game(D, &D.tab->hot, env, e);
  sf_store_env(D, env, e);
}


// ------------------------------------------------------------------------------------------------
// feature observations (ssf_env.py:95-157) from the state, one thread per env
// ------------------------------------------------------------------------------------------------
// computeExtra (game.cpp:282-312). The reference evaluates it inside updateShip while the ship is alive and keeps the
// values while it is dead; position, velocity and heading are frozen while dead, so evaluating it from the current
// state gives the same numbers. Before the first tick of a Game the reference's mExtra is uninitialised memory; the
// test harness zero-fills the Game, and so does this (tick 0 -> 0, 0, 0). fdist keeps the bug of the reference (Q11:
// the y term is ship.y - ship.y).
struct SfExtra { double vdir, aim, ndist; };
__device__ inline SfExtra sf_compute_extra(const SfEnv& e) {
  SfExtra x;
  x.vdir = 0.0; x.aim = 0.0; x.ndist = 0.0;
  if (e.q3.z == 0) return x;
  const double ang = (double)((unsigned)e.q0.x & SF_CORE_ANGLE_MASK);
  if (SF_DSQRT(SF_DADD(SF_DMUL(e.vel.x, e.vel.x), SF_DMUL(e.vel.y, e.vel.y))) != 0.0) {   // Vector::norm (vector.cpp:30-32)
    const double o = atan2(-SF_DSUB(SF_FORT_Y, e.pos.y), SF_DSUB(SF_FORT_X, e.pos.x));
    const double v = atan2(e.vel.y, e.vel.x);
    double d = SF_DSUB(v, o);
    if (d > SF_PI) d = SF_DSUB(d, SF_PI * 2);
    if (d < -SF_PI) d = SF_DADD(d, SF_PI * 2);
    x.vdir = sf_rad2deg(d);
  }
  double o = atan2(SF_DSUB(e.pos.y, SF_FORT_Y), SF_DSUB(e.pos.x, SF_FORT_X));
  o = SF_DADD(SF_DSUB(sf_rad2deg(o), ang), 180.0);
  if (o < -180.0) o = SF_DADD(o, 360.0);
  x.aim = o;
  const double dx = SF_DSUB(e.pos.x, SF_FORT_X), dy0 = SF_DSUB(e.pos.y, e.pos.y);
  const double fdist = SF_DSQRT(SF_DADD(SF_DMUL(dx, dx), SF_DMUL(dy0, dy0)));
  x.ndist = SF_DADD(-1.0, SF_DDIV(SF_DSUB(fdist, 40.0), SF_DDIV(SF_DSUB(200.0, 40.0), 2.0)));   // normDist (game.cpp:282-284)
  return x;
}
__device__ __forceinline__ double sf_clip1(double v) { return v < -1.0 ? -1.0 : (v > 1.0 ? 1.0 : v); }
__device__ __forceinline__ double sf_pymod360(double v) {  // Python's float `v % 360`
  double m = fmod(v, 360.0);
  if (m != 0.0) { if (m < 0.0) m = SF_DADD(m, 360.0); } else m = 0.0;
  return m;
}

template <class OutT>
__global__ void sf_features_kernel(SfDev D, int kind, OutT* out) {
  const int env = blockIdx.x * blockDim.x + threadIdx.x;
  if (env >= D.n) return;
  SfEnv e;
  sf_load_env(D, env, e);
  const SfExtra x = sf_compute_extra(e);
  const unsigned core = (unsigned)e.q0.x;
  const double alive = (core & SF_CORE_SHIP_ALIVE) ? 1.0 : 0.0, falive = (core & SF_CORE_FORT_ALIVE) ? 1.0 : 0.0;
  const int vuln = e.q1.z;
  const double kill_window = (vuln > 10 && e.q1.y < 250) ? 1.0 : 0.0;   // vulnerability_timer < fortressVulnerabilityTime
  const int nm = __popc((unsigned)e.q0.y & SF_PMASK_MISSILES), ns = __popc(((unsigned)e.q0.y >> SF_PMASK_SHELL_SHIFT) & 0xFu);
  const double fang = 10.0 * (double)((core >> SF_CORE_FANG_SHIFT) & 63u), sang = (double)(core & SF_CORE_ANGLE_MASK);
  const int nt = D.autoturn ? 2 : 4;
  double f[19];
  int nf;
  if (kind == SF_OBS_MONITORS) {          // ssf_env.py:96-108
    nf = 10;
    f[0] = nm > 0 ? 0.5 : -0.5; f[1] = falive != 0.0 ? 0.5 : -0.5; f[2] = vuln > 10 ? 0.5 : -0.5; f[3] = kill_window != 0.0 ? 0.5 : -0.5;
    f[4] = x.aim < 3 ? 0.5 : -0.5; f[5] = x.aim > 3 ? 0.5 : -0.5;
    f[6] = x.ndist > .75 ? 0.5 : -0.5; f[7] = x.ndist > .25 ? 0.5 : -0.5; f[8] = x.ndist < -.25 ? 0.5 : -0.5; f[9] = x.ndist < -.75 ? 0.5 : -0.5;
  } else if (kind == SF_OBS_NORMALIZED_FEATURES) {  // ssf_env.py:109-133 (sic: max(vulnerability, 10)); np.clip(f, -1, 1)
    nf = 15 + nt;
    f[0] = alive; f[1] = SF_DDIV(e.pos.x, 90.0); f[2] = SF_DDIV(e.pos.y, 92.0); f[3] = SF_DDIV(e.vel.x, 10.0); f[4] = SF_DDIV(e.vel.y, 10.0);
    f[5] = SF_DDIV(sang, 360.0); f[6] = SF_DDIV(x.aim, 180.0); f[7] = SF_DDIV(sf_pymod360(x.vdir), 360.0); f[8] = x.ndist;
    f[9] = falive; f[10] = SF_DDIV(fang, 360.0); f[11] = SF_DDIV((double)max(vuln, 10), 10.0); f[12] = kill_window;
    f[13] = SF_DDIV((double)nm, 20.0); f[14] = SF_DDIV((double)ns, 20.0);
    const double max_ticks = 5294.0;  // np.floor(180000 / 34), ssf_env.py:166
    f[15] = SF_DDIV((double)e.q2.x, max_ticks); f[16] = SF_DDIV((double)e.q2.y, max_ticks);
    f[17] = SF_DDIV((double)e.q2.z, max_ticks); f[18] = SF_DDIV((double)e.q2.w, max_ticks);
    for (int k = 0; k < 19; k++) f[k] = sf_clip1(f[k]);
  } else {                               // 'features', ssf_env.py:134-157
    nf = 15 + nt;
    f[0] = alive; f[1] = e.pos.x; f[2] = e.pos.y; f[3] = e.vel.x; f[4] = e.vel.y; f[5] = sang; f[6] = x.aim; f[7] = x.vdir; f[8] = x.ndist;
    f[9] = falive; f[10] = fang; f[11] = (double)vuln; f[12] = kill_window; f[13] = (double)nm; f[14] = (double)ns;
    f[15] = (double)e.q2.x; f[16] = (double)e.q2.y; f[17] = (double)e.q2.z; f[18] = (double)e.q2.w;
  }
  OutT* o = out + (size_t)env * nf;
  for (int k = 0; k < nf; k++) o[k] = (OutT)f[k];
}

// ---- state records (AoS <-> SoA), one thread per env ----
__global__ void sf_get_state_kernel(SfDev D, int first, int count, sf_state_record* out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int i = first + k;
  const SfTables* T = D.tab;
  SfEnv e;
  sf_load_env(D, i, e);
  sf_state_record r;
  memset(&r, 0, sizeof(r));
  unsigned core = (unsigned)e.q0.x;
  r.ship_x = e.pos.x; r.ship_y = e.pos.y; r.ship_vx = e.vel.x; r.ship_vy = e.vel.y;
  r.ship_angle = (double)(core & SF_CORE_ANGLE_MASK);
  r.fortress_angle = 10.0 * ((core >> SF_CORE_FANG_SHIFT) & 63u);
  r.fortress_last_angle = 10.0 * ((core >> SF_CORE_FLAST_SHIFT) & 63u);
  r.missile_mask = (unsigned)e.q0.y & SF_PMASK_MISSILES;
  r.shell_mask = ((unsigned)e.q0.y >> SF_PMASK_SHELL_SHIFT) & 0xFu;
  for (int s = 0; s < SF_MAX_MISSILES; s++) if ((r.missile_mask >> s) & 1) {
    double2 p = D.mpos[(size_t)s * D.n_pad + i];
    int a = D.mang[(size_t)s * D.n_pad + i];
    r.missile_x[s] = p.x; r.missile_y[s] = p.y; r.missile_angle[s] = a;
    r.missile_vx[s] = SF_DMUL(20.0, T->cos_deg[a]); r.missile_vy[s] = SF_DMUL(20.0, T->sin_deg[a]);
  }
  for (int s = 0; s < SF_DEV_SHELLS; s++) if ((r.shell_mask >> s) & 1) {
    double2 p = D.spos[(size_t)s * D.n_pad + i], v = D.svel[(size_t)s * D.n_pad + i];
    r.shell_x[s] = p.x; r.shell_y[s] = p.y; r.shell_vx[s] = v.x; r.shell_vy[s] = v.y;
    r.shell_angle[s] = D.sang[(size_t)s * D.n_pad + i];
  }
  r.points = __int_as_float(e.q3.x); r.raw_points = __int_as_float(e.q3.y);
  r.ship_alive = (core & SF_CORE_SHIP_ALIVE) != 0; r.fortress_alive = (core & SF_CORE_FORT_ALIVE) != 0;
  r.ship_death_timer = e.q0.z;
  r.fire_timer = e.q2.x; r.thrust_timer = e.q2.y; r.left_timer = e.q2.z; r.right_timer = e.q2.w;
  r.thrust_flag = (core & SF_CORE_THRUST) != 0; r.fire_flag = (core & SF_CORE_FIRE) != 0;
  r.left_flag = (core & SF_CORE_LEFT) != 0; r.right_flag = (core & SF_CORE_RIGHT) != 0;
  r.turn_flag = (r.left_flag && !r.right_flag) ? 1 : (!r.left_flag && r.right_flag) ? 2 : 0;  // game.cpp:265-270
  r.fortress_timer = e.q0.w; r.fortress_death_timer = e.q1.x; r.fortress_vuln_timer = e.q1.y;
  r.vulnerability = e.q1.z; r.tick = e.q3.z; r.time = e.q3.z * SF_TICK_MS;
  const int4 st0 = D.st0[i], st1 = D.st1[i], st2 = D.st2[i];
  r.stats[0] = st0.x; r.stats[1] = st0.y; r.stats[2] = st0.z; r.stats[3] = st0.w;
  r.stats[4] = st1.x; r.stats[5] = st1.y; r.stats[6] = st1.z; r.stats[7] = st1.w;
  r.stats[8] = st2.x; r.stats[9] = st2.y; r.stats[10] = st2.z; r.stats[11] = st2.w;
  r.stats[12] = e.st3.x;
  r.prev_vlner = e.q1.w;
  r.rng_seed = (unsigned)e.st3.w; r.rng_count = (unsigned)e.st3.z;
  r.ep_return = e.q3.w;
  out[k] = r;
}

__global__ void sf_set_state_kernel(SfDev D, int first, int count, const sf_state_record* in) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= count) return;
  const int i = first + k;
  const sf_state_record& r = in[k];
  SfEnv e;
  sf_load_env(D, i, e);
  // the rand stream is re-derived from (seed, count)
  if ((unsigned)e.st3.w != r.rng_seed || (unsigned)e.st3.z > r.rng_count) sf_srand(D, i, e, r.rng_seed);
  while ((unsigned)e.st3.z < r.rng_count) (void)sf_rand(D, i, e);
  unsigned core = ((unsigned)(int)r.ship_angle & SF_CORE_ANGLE_MASK) | (((unsigned)(int)(r.fortress_angle / 10.0) & 63u) << SF_CORE_FANG_SHIFT) |
                  (((unsigned)(int)(r.fortress_last_angle / 10.0) & 63u) << SF_CORE_FLAST_SHIFT);
  if (r.ship_alive) core |= SF_CORE_SHIP_ALIVE;
  if (r.fortress_alive) core |= SF_CORE_FORT_ALIVE;
  if (r.fire_flag) core |= SF_CORE_FIRE;
  if (r.thrust_flag) core |= SF_CORE_THRUST;
  if (r.left_flag) core |= SF_CORE_LEFT;
  if (r.right_flag) core |= SF_CORE_RIGHT;
  e.pos = make_double2(r.ship_x, r.ship_y); e.vel = make_double2(r.ship_vx, r.ship_vy);
  e.q0 = make_int4((int)core, (int)((r.missile_mask & SF_PMASK_MISSILES) | ((r.shell_mask & 0xFu) << SF_PMASK_SHELL_SHIFT)), r.ship_death_timer, r.fortress_timer);
  e.q1 = make_int4(r.fortress_death_timer, r.fortress_vuln_timer, r.vulnerability, r.prev_vlner);
  e.q2 = make_int4(r.fire_timer, r.thrust_timer, r.left_timer, r.right_timer);
  e.q3 = make_int4(__float_as_int(r.points), __float_as_int(r.raw_points), r.tick, r.ep_return);
  D.st0[i] = make_int4(r.stats[0], r.stats[1], r.stats[2], r.stats[3]);
  D.st1[i] = make_int4(r.stats[4], r.stats[5], r.stats[6], r.stats[7]);
  D.st2[i] = make_int4(r.stats[8], r.stats[9], r.stats[10], r.stats[11]);
  e.st3.x = r.stats[12];
  for (int s = 0; s < SF_MAX_MISSILES; s++) if ((r.missile_mask >> s) & 1) {
    D.mpos[(size_t)s * D.n_pad + i] = make_double2(r.missile_x[s], r.missile_y[s]);
    D.mang[(size_t)s * D.n_pad + i] = (short)(int)r.missile_angle[s];
  }
  for (int s = 0; s < SF_DEV_SHELLS; s++) if ((r.shell_mask >> s) & 1) {
    D.spos[(size_t)s * D.n_pad + i] = make_double2(r.shell_x[s], r.shell_y[s]);
    D.svel[(size_t)s * D.n_pad + i] = make_double2(r.shell_vx[s], r.shell_vy[s]);
    D.sang[(size_t)s * D.n_pad + i] = r.shell_angle[s];
  }
  D.expstamp[i] = 0;  // a forced state may place a dead ship anywhere
  D.expo_meta[i] = make_uint2(0u, 0u);
  sf_store_env(D, i, e);
}

// ================================================================================================
// C-ABI
// ================================================================================================
static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(SF_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

#define SF_MAX_MIRRORS 4
#define SF_HOST_MAX_SLICES 8
struct sf_handle {
  SfDev dev;
  int device;
  int action_set;
  int gametype;  // 0 youturn 1 autoturn 2 test-youturn 3 test-autoturn
  int num_sms;
  void* slab;
  size_t slab_bytes;
  SfTables* h_tab;
  // pinned staging for sf_step_host
  int* d_actions; unsigned char* d_obs; int* d_reward; unsigned char* d_done; unsigned char* d_kill; unsigned* d_events;
  size_t staging_obs_bytes;
  SfSched* d_sched;  // sf_rollout_kernel's scheduling state (group hand-out, cost-dealt envs)
  cudaStream_t host_compute, host_copy;  // sf_step_host: kernels of slice k + 1 overlap the device->host copy of slice k
  cudaEvent_t host_ev[SF_HOST_MAX_SLICES];
  int host_slices;
  // SF_FLAG_HOST_DELTA: device copy of what the caller's page-locked h_obs holds, so that only changed bytes cross PCIe
  struct Mirror { const void* host; unsigned char* d; size_t cap, bytes; unsigned long long used; };
  Mirror mirrors[SF_MAX_MIRRORS];     // one per host buffer (a caller may rotate a few buffers); least recently used goes first
  unsigned long long mirror_clock;
  unsigned long long* d_delta_stats;  // [0] observation bytes written to the host by delta calls, [1] delta calls
  unsigned long long full_calls;      // calls that sent whole frames (first call, new buffer, flag absent)
  int delta_lanes;                    // granule of the delta updates in 16-byte lanes (4; SF_DELTA_GRANULE = 16 ... 256 bytes overrides)
};

extern "C" const char* sf_last_error(void) { return g_err.c_str(); }
extern "C" int sf_version(void) { return 100; }

static int gametype_of(const char* s) {
  if (!s) return -1;
  if (!strcmp(s, "youturn")) return 0;
  if (!strcmp(s, "autoturn")) return 1;
  if (!strcmp(s, "test-youturn")) return 2;
  if (!strcmp(s, "test-autoturn")) return 3;
  return -1;
}

// P2: action tables (ssf_env.py:65-90)
static int build_actions(int gametype, int action_set, int* km) {
  bool youturn = gametype == 0 || gametype == 2;
  if (action_set == 1) {
    const int tab[5] = {0, SF_KEY_FIRE, SF_KEY_THRUST, SF_KEY_LEFT, SF_KEY_RIGHT};
    int n = youturn ? 5 : 3;
    for (int i = 0; i < n; i++) km[i] = tab[i];
    return n;
  }
  if (action_set == 0 && !youturn) {  // np.meshgrid([0,1],[0,1]).T.reshape(-1,2): fire = bit 1, thrust = bit 0
    for (int a = 0; a < 4; a++) km[a] = ((a >> 1) & 1 ? SF_KEY_FIRE : 0) | ((a & 1) ? SF_KEY_THRUST : 0);
    return 4;
  }
  if (action_set == 0 || action_set == -1) {  // 4-key meshgrid: row = right*8 + left*4 + fire*2 + thrust
    for (int a = 0; a < 16; a++) {
      int m = ((a >> 1) & 1 ? SF_KEY_FIRE : 0) | ((a & 1) ? SF_KEY_THRUST : 0) | ((a >> 2) & 1 ? SF_KEY_LEFT : 0) | ((a >> 3) & 1 ? SF_KEY_RIGHT : 0);
      km[a] = youturn ? m : (m & (SF_KEY_FIRE | SF_KEY_THRUST));
    }
    return 16;
  }
  return -1;
}

template <class T>
static T* carve(char*& p, size_t count) {
  T* r = reinterpret_cast<T*>(p);
  size_t bytes = (count * sizeof(T) + 255) & ~(size_t)255;
  p += bytes;
  return r;
}

static size_t layout(SfDev& d, char* base) {
  char* p = base;
  size_t np = (size_t)d.n_pad;
  d.pos = carve<double2>(p, np); d.vel = carve<double2>(p, np);
  d.q0 = carve<int4>(p, np); d.q1 = carve<int4>(p, np); d.q2 = carve<int4>(p, np); d.q3 = carve<int4>(p, np);
  d.st0 = carve<int4>(p, np); d.st1 = carve<int4>(p, np); d.st2 = carve<int4>(p, np); d.st3 = carve<int4>(p, np);
  d.mpos = carve<double2>(p, np * SF_MAX_MISSILES); d.mang = carve<short>(p, np * SF_MAX_MISSILES);
  d.spos = carve<double2>(p, np * SF_DEV_SHELLS); d.svel = carve<double2>(p, np * SF_DEV_SHELLS); d.sang = carve<double>(p, np * SF_DEV_SHELLS);
  d.rng = carve<unsigned>(p, np * SF_RNG_WORDS);
  d.expc = carve<unsigned char>(p, np * SF_EXP_W * SF_EXP_W);
  d.expstamp = carve<unsigned>(p, np);
  d.expo = carve<unsigned char>(p, np * SF_EXPO_BYTES);
  d.expo_meta = carve<uint2>(p, np);
  d.epi = carve<unsigned long long>(p, SF_NUM_EPISODE_STATS);
  d.tab = carve<SfTables>(p, 1);
  d.static_image = carve<unsigned char>(p, SF_STATIC_BYTES);
  return (size_t)(p - base);
}

extern "C" int sf_destroy(sf_handle* h);

extern "C" int sf_create(const char* gametype, int action_set, int n_envs, int device, sf_handle** out) {
  if (!out) return fail(SF_ERR_INVALID, "out is NULL");
  *out = nullptr;
  int gt = gametype_of(gametype);
  if (gt < 0) return fail(SF_ERR_INVALID, std::string("cannot initialize Game. Unknown config value: `") + (gametype ? gametype : "(null)") + "'");
  if (n_envs <= 0) return fail(SF_ERR_INVALID, "n_envs must be positive");
  int km[16];
  int na = build_actions(gt, action_set, km);
  if (na < 0) return fail(SF_ERR_INVALID, "action_set must be 1, 0 or -1");
  int ndev = 0;
  CUDA_TRY(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(SF_ERR_CUDA, "no such CUDA device (this library has no CPU fallback)");
  CUDA_TRY(cudaSetDevice(device));
  int num_sms = 0;
  CUDA_TRY(cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, device));

  sf_handle* h = new sf_handle();
  memset(h, 0, sizeof(*h));
  h->device = device; h->num_sms = num_sms > 0 ? num_sms : 148; h->action_set = action_set; h->gametype = gt;
  SfDev& d = h->dev;
  d.n = n_envs; d.n_pad = (n_envs + 31) & ~31;
  d.autoturn = (gt == 1 || gt == 3);
  bool test = gt >= 2;
  d.shaped = !test;                      // ssf_env.py:235
  d.destroy_fortress = test ? 100 : 1;   // configs.cpp:8,57,68
  d.death_penalty = test ? 100 : 1;      // configs.cpp:9,58,69
  d.missile_penalty = test ? 2.0f : (float)0.05;  // configs.cpp:10,59,70 (double -> float parameter, game.cpp:104)
  d.num_actions = na;
  for (int i = 0; i < 16; i++) d.keymask_of_action[i] = i < na ? km[i] : 0;
  d.first_global_env = 0;

  h->h_tab = new SfTables();
  char err[256] = {0};
  if (sf_build_tables(h->h_tab, err, sizeof(err))) { std::string m = err; delete h->h_tab; delete h; return fail(SF_ERR_INVALID, "table build failed: " + m); }

  // every failure below releases what has been allocated (sf_destroy frees whatever is non-NULL)
  auto bail = [&](const char* what, cudaError_t ce) { std::string m = std::string(what) + ": " + cudaGetErrorString(ce); sf_destroy(h); return fail(SF_ERR_CUDA, m); };
  cudaError_t ce;
  h->slab_bytes = layout(d, nullptr);
  if ((ce = cudaMalloc(&h->slab, h->slab_bytes)) != cudaSuccess) return bail("cudaMalloc state slab", ce);
  layout(d, (char*)h->slab);
  if ((ce = cudaMalloc(&h->d_sched, sizeof(SfSched))) != cudaSuccess) return bail("cudaMalloc scheduler state", ce);
  if ((ce = cudaMemset(h->d_sched, 0, sizeof(SfSched))) != cudaSuccess) return bail("cudaMemset scheduler state", ce);
  if ((ce = cudaMemset(h->slab, 0, h->slab_bytes)) != cudaSuccess) return bail("cudaMemset state slab", ce);
  if ((ce = cudaMemcpy((void*)d.tab, h->h_tab, sizeof(SfTables), cudaMemcpyHostToDevice)) != cudaSuccess) return bail("upload tables", ce);
  if ((ce = cudaFuncSetAttribute(sf_rollout_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK))) != cudaSuccess) return bail("shared memory opt-in (rollout)", ce);
  if ((ce = cudaFuncSetAttribute(sf_rollout_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK))) != cudaSuccess) return bail("shared memory opt-in (rollout, host delta)", ce);
  if ((ce = cudaFuncSetAttribute(sf_render_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK))) != cudaSuccess) return bail("shared memory opt-in (render)", ce);
  if ((ce = cudaFuncSetAttribute(sf_pack_static_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK))) != cudaSuccess) return bail("shared memory opt-in (pack)", ce);
  sf_pack_static_kernel<<<1, SF_BLOCK, SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK)>>>(d, const_cast<unsigned char*>(d.static_image));
  if ((ce = cudaGetLastError()) != cudaSuccess) return bail("static image kernel launch", ce);
  // default seeding: every env replays srand(1) — the reference never seeds libc (game.cpp:137-148)
  sf_seed_kernel<<<(d.n + 127) / 128, 128>>>(d, nullptr);
  if ((ce = cudaGetLastError()) != cudaSuccess) return bail("seed kernel launch", ce);
  if ((ce = cudaDeviceSynchronize()) != cudaSuccess) return bail("seed kernel", ce);
  *out = h;
  return SF_OK;
}

extern "C" int sf_destroy(sf_handle* h) {
  if (!h) return SF_OK;
  cudaSetDevice(h->device);
  if (h->slab) cudaFree(h->slab);
  if (h->d_actions) cudaFree(h->d_actions);
  if (h->d_obs) cudaFree(h->d_obs);
  if (h->d_reward) cudaFree(h->d_reward);
  if (h->d_done) cudaFree(h->d_done);
  if (h->d_kill) cudaFree(h->d_kill);
  if (h->d_events) cudaFree(h->d_events);
  if (h->d_sched) cudaFree(h->d_sched);
  for (auto& m : h->mirrors) if (m.d) cudaFree(m.d);
  if (h->d_delta_stats) cudaFree(h->d_delta_stats);
  if (h->host_compute) cudaStreamDestroy(h->host_compute);
  if (h->host_copy) cudaStreamDestroy(h->host_copy);
  for (int k = 0; k < SF_HOST_MAX_SLICES; k++) if (h->host_ev[k]) cudaEventDestroy(h->host_ev[k]);
  delete h->h_tab;
  delete h;
  return SF_OK;
}

extern "C" int sf_num_envs(const sf_handle* h) { return h ? h->dev.n : -1; }
extern "C" int sf_num_actions(const sf_handle* h) { return h ? h->dev.num_actions : -1; }
extern "C" int sf_action_keymask(const sf_handle* h, int action) {
  if (!h || action < 0 || action >= h->dev.num_actions) return -1;
  return h->dev.keymask_of_action[action];
}
extern "C" long long sf_state_bytes(const sf_handle* h) { return h ? (long long)h->slab_bytes : -1; }

extern "C" int sf_seed(sf_handle* h, const uint32_t* h_seeds, long long first_global_env, void* stream) {
  if (!h) return fail(SF_ERR_INVALID, "handle is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  h->dev.first_global_env = first_global_env;
  unsigned* d_seeds = nullptr;
  cudaError_t ce = cudaSuccess;
  if (h_seeds) {
    CUDA_TRY(cudaMalloc(&d_seeds, sizeof(unsigned) * h->dev.n));
    ce = cudaMemcpyAsync(d_seeds, h_seeds, sizeof(unsigned) * h->dev.n, cudaMemcpyHostToDevice, st);
  }
  if (ce == cudaSuccess) {
    sf_seed_kernel<<<(h->dev.n + 127) / 128, 128, 0, st>>>(h->dev, d_seeds);
    ce = cudaGetLastError();
  }
  if (d_seeds) {  // the temporary is released on every path
    cudaError_t cs = cudaStreamSynchronize(st);
    if (ce == cudaSuccess) ce = cs;
    cudaFree(d_seeds);
  }
  if (ce != cudaSuccess) return fail(SF_ERR_CUDA, std::string("sf_seed: ") + cudaGetErrorString(ce));
  return SF_OK;
}

// Envs per group (= per block and tick). While every group can have its own SM the envs are spread evenly over
// the SMs (4096 envs -> 28 per block, 147 blocks); beyond that a group is a full warp of 32 stepping lanes and
// the blocks are persistent over their groups.
static void group_shape(const sf_handle* h, int n, int* EB, int* ngroups, int* blocks) {
  const int sms = h->num_sms * SF_BLOCKS_PER_SM, teams = sms;
  int eb = n <= teams * SF_GROUP_ENVS ? (n + teams - 1) / teams : SF_GROUP_ENVS;
  if (const char* ov = getenv("SF_ENVS_PER_BLOCK")) { int e = atoi(ov); if (e >= 1 && e <= SF_GROUP_ENVS) eb = e; }  // tuning knob
  *EB = eb;
  *ngroups = (n + eb - 1) / eb;
  *blocks = std::min(*ngroups, sms);
}

static int launch_render(sf_handle* h, unsigned char* d_obs, int flags, const unsigned char* d_mask, cudaStream_t st) {
  int EB, ngroups, blocks;
  group_shape(h, h->dev.n, &EB, &ngroups, &blocks);
  sf_render_kernel<<<blocks, SF_BLOCK, SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK), st>>>(h->dev, d_obs, flags, d_mask, EB, ngroups);
  CUDA_TRY(cudaGetLastError());
  return SF_OK;
}

// ------------------------------------------------------------------------------------------------
// policy input (consumer side of the on-device rollout): 4-frame stack -> space-to-depth bf16 NHWC
// ------------------------------------------------------------------------------------------------
#include <cuda_bf16.h>
// One block per env. For each of the 21 output rows Y the 4 frames x 4 pixel rows it needs (1344 B) are staged in
// shared memory with coalesced 32-bit loads (two tiles: one barrier per row), then the 21 x 64 bf16 of the row go out
// as coalesced 16-byte stores (8 channels = frame f, rows dy0, dy0 + 1, 4 columns of block X). u8 -> bf16(u8 / 255)
// through a 256-entry table built with the exact formula (u8 -> fp32, / 255 in fp32, round to bf16, like torch).
#define SF_PI_THREADS 192
template <bool F32>
__global__ void __launch_bounds__(SF_PI_THREADS) sf_policy_input_kernel(const unsigned char* __restrict__ frames, long long fstride, int n, const int* __restrict__ valid,
                                                                       uint4* __restrict__ out) {
  __shared__ unsigned tile[2][4 * 4 * 21];  // [frame][dy][X] : 4 pixels each
  __shared__ unsigned lut[256];             // bf16 bits of u8 / 255 (F32: the fp32 bits)
  const int env = blockIdx.x, tid = threadIdx.x;
  for (int v = tid; v < 256; v += SF_PI_THREADS) {
    // fp32: torch's CUDA `x / 255.0` multiplies by the fp32 reciprocal of the scalar; the table reproduces that bit for bit
    if (F32) lut[v] = __float_as_uint(__fmul_rn((float)v, __fdiv_rn(1.0f, 255.0f)));
    else { const __nv_bfloat16 h = __float2bfloat16_rn(__fdiv_rn((float)v, 255.0f)); lut[v] = *reinterpret_cast<const unsigned short*>(&h); }
  }
  const int nvalid = min(max(valid[env], 0), 4);
  // what this thread loads (words k0 = tid and k1 = tid + 192 of the 336-word tile) and stores (chunk tid < 168)
  const unsigned char* base = frames + (long long)env * (84 * 84);
  const int k1 = tid + SF_PI_THREADS;
  const int f0 = tid / 84, r0 = tid - f0 * 84, f1 = k1 / 84, r1 = k1 - f1 * 84;
  const bool on0 = f0 >= 4 - nvalid, on1 = k1 < 336 && f1 >= 4 - nvalid;
  const unsigned char* p0 = base + (long long)f0 * fstride + (r0 / 21) * 84 + (r0 % 21) * 4;
  const unsigned char* p1 = base + (long long)(k1 < 336 ? f1 : 0) * fstride + (r1 / 21) * 84 + (r1 % 21) * 4;
  const int X = tid >> 3, c8 = tid & 7, f = c8 >> 1, dy0 = (c8 & 1) * 2;
  const int ta = (f * 4 + dy0) * 21 + X, tb = ta + 21;
  uint4* orow = out + (long long)env * (21 * 21 * 8) * (F32 ? 2 : 1);
  auto cvt2 = [&](unsigned lo, unsigned hi) { return lut[lo] | (lut[hi] << 16); };
#pragma unroll 1
  for (int Y = 0; Y < 21; Y++) {
    unsigned* T = tile[Y & 1];
    T[tid] = on0 ? *reinterpret_cast<const unsigned*>(p0 + Y * (4 * 84)) : 0u;
    if (k1 < 336) T[k1] = on1 ? *reinterpret_cast<const unsigned*>(p1 + Y * (4 * 84)) : 0u;
    __syncthreads();  // (the other tile is free again: every thread passed the previous barrier after reading it)
    if (tid < 21 * 8) {
      const unsigned a = T[ta], b = T[tb];
      if (F32) {
        uint4 o0, o1;
        o0.x = lut[a & 255u]; o0.y = lut[(a >> 8) & 255u]; o0.z = lut[(a >> 16) & 255u]; o0.w = lut[a >> 24];
        o1.x = lut[b & 255u]; o1.y = lut[(b >> 8) & 255u]; o1.z = lut[(b >> 16) & 255u]; o1.w = lut[b >> 24];
        orow[(Y * (21 * 8) + tid) * 2] = o0; orow[(Y * (21 * 8) + tid) * 2 + 1] = o1;
      } else {
        uint4 o;
        o.x = cvt2(a & 255u, (a >> 8) & 255u); o.y = cvt2((a >> 16) & 255u, a >> 24);
        o.z = cvt2(b & 255u, (b >> 8) & 255u); o.w = cvt2((b >> 16) & 255u, b >> 24);
        orow[Y * (21 * 8) + tid] = o;
      }
    }
  }
}

static int policy_input_common(const uint8_t* d_frames, long long frame_stride_bytes, int n, const int32_t* d_valid, void* d_out, bool f32, void* stream) {
  if (!d_frames || !d_valid || !d_out || n <= 0 || (frame_stride_bytes & 3)) return fail(SF_ERR_INVALID, "sf_policy_input: bad arguments");
  cudaPointerAttributes pa;  // no handle: launch on the device that owns the frames, whatever the caller's current device is
  CUDA_TRY(cudaPointerGetAttributes(&pa, d_frames));
  if (pa.type != cudaMemoryTypeDevice && pa.type != cudaMemoryTypeManaged) return fail(SF_ERR_INVALID, "sf_policy_input: d_frames is not device memory");
  CUDA_TRY(cudaSetDevice(pa.device));
  if (f32) sf_policy_input_kernel<true><<<(unsigned)n, SF_PI_THREADS, 0, (cudaStream_t)stream>>>(d_frames, frame_stride_bytes, n, d_valid, reinterpret_cast<uint4*>(d_out));
  else sf_policy_input_kernel<false><<<(unsigned)n, SF_PI_THREADS, 0, (cudaStream_t)stream>>>(d_frames, frame_stride_bytes, n, d_valid, reinterpret_cast<uint4*>(d_out));
  CUDA_TRY(cudaGetLastError());
  return SF_OK;
}
extern "C" int sf_policy_input(const uint8_t* d_frames, long long frame_stride_bytes, int n, const int32_t* d_valid, void* d_out_bf16, void* stream) {
  return policy_input_common(d_frames, frame_stride_bytes, n, d_valid, d_out_bf16, false, stream);
}
extern "C" int sf_policy_input_f32(const uint8_t* d_frames, long long frame_stride_bytes, int n, const int32_t* d_valid, float* d_out_f32, void* stream) {
  return policy_input_common(d_frames, frame_stride_bytes, n, d_valid, d_out_f32, true, stream);
}

extern "C" int sf_num_features(const sf_handle* h, int obs_type) {
  if (!h) return -1;
  if (obs_type == SF_OBS_MONITORS) return 10;
  if (obs_type == SF_OBS_FEATURES || obs_type == SF_OBS_NORMALIZED_FEATURES) return 15 + (h->dev.autoturn ? 2 : 4);
  return -1;
}
static int features_common(sf_handle* h, int obs_type, void* d_out, bool f64, void* stream) {
  if (!h || !d_out) return fail(SF_ERR_INVALID, "handle or out is NULL");
  if (sf_num_features(h, obs_type) < 0) return fail(SF_ERR_INVALID, "obs_type must be SF_OBS_FEATURES, SF_OBS_NORMALIZED_FEATURES or SF_OBS_MONITORS");
  CUDA_TRY(cudaSetDevice(h->device));
  if (f64) sf_features_kernel<double><<<(h->dev.n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->dev, obs_type, (double*)d_out);
  else sf_features_kernel<float><<<(h->dev.n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(h->dev, obs_type, (float*)d_out);
  CUDA_TRY(cudaGetLastError());
  return SF_OK;
}
extern "C" int sf_features(sf_handle* h, int obs_type, float* d_out, void* stream) { return features_common(h, obs_type, d_out, false, stream); }
extern "C" int sf_features_f64(sf_handle* h, int obs_type, double* d_out, void* stream) { return features_common(h, obs_type, d_out, true, stream); }

extern "C" int sf_set_ticks(sf_handle* h, const int32_t* h_ticks) {
  if (!h || !h_ticks) return fail(SF_ERR_INVALID, "handle or ticks is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  int* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(int) * (size_t)h->dev.n));
  cudaError_t ce = cudaDeviceSynchronize();  // synchronous call: ordered after whatever any stream has queued for this slab
  if (ce == cudaSuccess) ce = cudaMemcpy(d, h_ticks, sizeof(int) * (size_t)h->dev.n, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) {
    sf_set_ticks_kernel<<<(h->dev.n + 127) / 128, 128>>>(h->dev, d);
    ce = cudaDeviceSynchronize();
  }
  cudaFree(d);
  if (ce != cudaSuccess) return fail(SF_ERR_CUDA, std::string("sf_set_ticks: ") + cudaGetErrorString(ce));
  return SF_OK;
}

extern "C" int sf_reset(sf_handle* h, const uint8_t* d_mask, int clear_prev_vlner, uint8_t* d_obs, int flags, void* stream) {
  if (!h) return fail(SF_ERR_INVALID, "handle is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)stream;
  sf_reset_kernel<<<(h->dev.n + 127) / 128, 128, 0, st>>>(h->dev, d_mask, clear_prev_vlner);
  CUDA_TRY(cudaGetLastError());
  if (d_obs) return launch_render(h, d_obs, flags, d_mask, st);
  return SF_OK;
}

extern "C" int sf_render(sf_handle* h, uint8_t* d_obs, int flags, void* stream) {
  if (!h || !d_obs) return fail(SF_ERR_INVALID, "handle or obs is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  return launch_render(h, d_obs, flags, nullptr, (cudaStream_t)stream);
}

static int launch_rollout(sf_handle* h, const SfRollArgs& a, cudaStream_t st) {
  const SfDev& d = h->dev;
  bool render = (a.flags & SF_FLAG_RENDER) && a.obs;
  if (render) {
    SfRollArgs b = a;
    int blocks;
    group_shape(h, b.envn, &b.EB, &b.ngroups, &blocks);
    b.sched = (a.env0 == 0 && a.envn == d.n && !a.host_obs) ? h->d_sched : nullptr;  // slices of the slab (sf_step_host), in-kernel delta: static groups
    if (b.host_obs) sf_rollout_kernel<true><<<blocks, SF_BLOCK, SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK), st>>>(d, b);
    else sf_rollout_kernel<false><<<blocks, SF_BLOCK, SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK), st>>>(d, b);
  } else {
    sf_step_only_kernel<<<(a.envn + 127) / 128, 128, 0, st>>>(d, a);
  }
  CUDA_TRY(cudaGetLastError());
  return SF_OK;
}

extern "C" int sf_step(sf_handle* h, const int32_t* d_actions, uint8_t* d_obs, int32_t* d_reward, uint8_t* d_done,
                       uint8_t* d_fortkill, uint32_t* d_events, int flags, void* stream) {
  if (!h || !d_actions) return fail(SF_ERR_INVALID, "handle or actions is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  SfRollArgs a = {};
  a.T = 1; a.EB = 1; a.ngroups = 0; a.flags = flags; a.actions = d_actions; a.action_seed = 0; a.t0 = 0;
  a.obs = d_obs; a.reward = d_reward; a.done = d_done; a.fortkill = d_fortkill; a.events = d_events; a.sched = nullptr; a.env0 = 0; a.envn = h->dev.n;
  return launch_rollout(h, a, (cudaStream_t)stream);
}

extern "C" int sf_rollout(sf_handle* h, int T, const int32_t* d_actions, uint32_t action_seed, long long t0, uint8_t* d_obs,
                          int32_t* d_reward, uint8_t* d_done, uint8_t* d_fortkill, int flags, void* stream) {
  if (!h || T <= 0) return fail(SF_ERR_INVALID, "handle is NULL or T <= 0");
  CUDA_TRY(cudaSetDevice(h->device));
  SfRollArgs a = {};
  a.T = T; a.EB = 1; a.ngroups = 0; a.flags = flags; a.actions = d_actions; a.action_seed = action_seed; a.t0 = t0;
  a.obs = d_obs; a.reward = d_reward; a.done = d_done; a.fortkill = d_fortkill; a.events = nullptr; a.sched = nullptr; a.env0 = 0; a.envn = h->dev.n;
  return launch_rollout(h, a, (cudaStream_t)stream);
}

extern "C" int sf_synthetic_action(uint32_t action_seed, long long global_env, long long t, int num_actions) {
  return sf_hash_action(action_seed, (unsigned long long)global_env, (unsigned long long)t, num_actions);
}

// SF_FLAG_HOST_DELTA. Between two steps of an env only a few dozen bytes of its frame change (the moving objects and
// now and then a score digit), so the frames are not copied: `mirror` is the device's copy of what the caller's
// page-locked buffer holds, and this kernel compares the new frames with it 16 bytes per thread and stores, straight
// into the host buffer (its device alias: posted writes across PCIe), the granules (64 bytes: a host cache line) that differ. The host buffer
// ends up identical to a full copy; the traffic is ~600-1100 B per env-step instead of 7056.
template <int LANES>  // lanes (of 16 bytes) per granule
__global__ void __launch_bounds__(256) sf_host_delta_kernel(const uint4* __restrict__ cur, uint4* __restrict__ mirror, uint4* __restrict__ host,
                                                             size_t n16, int tail, unsigned long long* stats) {
  __shared__ unsigned block_sent;
  if (threadIdx.x == 0) block_sent = 0;
  __syncthreads();
  const unsigned lane = threadIdx.x & 31;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  const size_t n16_warp = (n16 + 31) & ~(size_t)31;  // whole warps stay in the loop for the ballot
  unsigned sent = 0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16_warp; i += stride) {
    const bool in = i < n16;
    uint4 a = make_uint4(0, 0, 0, 0), b = a;
    if (in) { a = __ldcg(cur + i); b = __ldcg(mirror + i); }
    const bool diff = (a.x != b.x) | (a.y != b.y) | (a.z != b.z) | (a.w != b.w);
    const unsigned m = __ballot_sync(0xffffffffu, diff);
    if (m == 0) continue;
    if (in && ((m >> (lane & ~(unsigned)(LANES - 1))) & ((1u << LANES) - 1u))) { host[i] = a; sent++; }  // whole granules
    if (diff) mirror[i] = a;
  }
  if (tail && blockIdx.x == 0 && threadIdx.x < (unsigned)tail) {  // the last (bytes % 16) bytes go every time
    const size_t o = n16 * 16 + threadIdx.x;
    const unsigned char v = reinterpret_cast<const unsigned char*>(cur)[o];
    reinterpret_cast<unsigned char*>(host)[o] = v; reinterpret_cast<unsigned char*>(mirror)[o] = v;
  }
  for (int o = 16; o; o >>= 1) sent += __shfl_xor_sync(0xffffffffu, sent, o);
  if (lane == 0 && sent) atomicAdd(&block_sent, sent);
  __syncthreads();
  if (threadIdx.x == 0) {
    if (block_sent) atomicAdd(&stats[0], (unsigned long long)block_sent * 16u);
    if (blockIdx.x == 0) atomicAdd(&stats[1], 1ull);
  }
}

static int ensure_staging(sf_handle* h, size_t obs_bytes) {
  size_t n = (size_t)h->dev.n;
  if (!h->host_compute) {
    CUDA_TRY(cudaStreamCreateWithFlags(&h->host_compute, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->host_copy, cudaStreamNonBlocking));
    for (int k = 0; k < SF_HOST_MAX_SLICES; k++) CUDA_TRY(cudaEventCreateWithFlags(&h->host_ev[k], cudaEventDisableTiming));
    h->host_slices = 4;
    if (const char* ov = getenv("SF_HOST_SLICES")) { int v = atoi(ov); if (v >= 1 && v <= SF_HOST_MAX_SLICES) h->host_slices = v; }  // tuning knob
    h->delta_lanes = 4;  // 64 bytes: whole host cache lines (no partial-line writes in the host's memory system)
    if (const char* ov = getenv("SF_DELTA_GRANULE")) { int v = atoi(ov); if (v == 16 || v == 32 || v == 64 || v == 128 || v == 256) h->delta_lanes = v / 16; }  // tuning knob
  }
  if (!h->d_actions) {
    CUDA_TRY(cudaMalloc(&h->d_actions, n * 4)); CUDA_TRY(cudaMalloc(&h->d_reward, n * 4));
    CUDA_TRY(cudaMalloc(&h->d_done, n)); CUDA_TRY(cudaMalloc(&h->d_kill, n)); CUDA_TRY(cudaMalloc(&h->d_events, n * 4));
  }
  if (obs_bytes > h->staging_obs_bytes) {
    if (h->d_obs) { cudaFree(h->d_obs); h->d_obs = nullptr; }
    CUDA_TRY(cudaMalloc(&h->d_obs, obs_bytes));
    h->staging_obs_bytes = obs_bytes;
  }
  return SF_OK;
}

extern "C" int sf_host_alloc(void** out, long long bytes) {
  if (!out || bytes <= 0) return fail(SF_ERR_INVALID, "sf_host_alloc: bad arguments");
  CUDA_TRY(cudaHostAlloc(out, (size_t)bytes, cudaHostAllocPortable | cudaHostAllocMapped));
  return SF_OK;
}
extern "C" int sf_host_free(void* p) {
  if (p) CUDA_TRY(cudaFreeHost(p));
  return SF_OK;
}

// The device's address for a page-locked host buffer (NULL for pageable memory): kernels read / write such buffers in
// place across PCIe, which for the small per-env arrays of a step replaces a DMA copy (several microseconds each).
static void* device_alias(const void* host) {
  cudaPointerAttributes at;
  if (!host || cudaPointerGetAttributes(&at, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return at.type == cudaMemoryTypeHost ? at.devicePointer : nullptr;
}

#ifdef SF_HOST_PROFILE  // tools/: host-side phases of sf_step_host (ns, averaged, printed every 512 calls)
#include <time.h>
static double g_hp[8]; static long g_hp_calls; static double g_hp_last;
static double hp_now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e9 + ts.tv_nsec; }
#define SF_HP(k) do { double t_ = hp_now(); if (k == 0) { g_hp_calls++; } else g_hp[k] += t_ - g_hp_last; g_hp_last = t_; \
  if (k == 5 && g_hp_calls % 512 == 0) { fprintf(stderr, "sf_step_host phases (us): entry %.1f legacy-sync %.1f aliases %.1f launches %.1f wait %.1f\n", \
    g_hp[1] / 512e3, g_hp[2] / 512e3, g_hp[3] / 512e3, g_hp[4] / 512e3, g_hp[5] / 512e3); for (double& v : g_hp) v = 0; } } while (0)
#else
#define SF_HP(k) do {} while (0)
#endif

extern "C" int sf_step_host(sf_handle* h, const int32_t* h_actions, uint8_t* h_obs, int32_t* h_reward, uint8_t* h_done,
                            uint8_t* h_fortkill, uint32_t* h_events, int flags) {
  SF_HP(0);
  if (!h || !h_actions) return fail(SF_ERR_INVALID, "handle or actions is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  size_t n = (size_t)h->dev.n;
  size_t per = (flags & SF_FLAG_NATIVE_OBS) ? (size_t)SF_NAT_H * SF_NAT_W : (size_t)84 * 84;
  bool render = (flags & SF_FLAG_RENDER) && h_obs;
  int rc = ensure_staging(h, render ? n * per : 0);
  if (rc) return rc;
  // Synchronous, on the handle's own streams. Whole-frame mode: the slab is stepped in slices of consecutive envs, the
  // kernel of slice k + 1 runs while the frames of slice k cross PCIe (one cudaMemcpyAsync per buffer and slice), so only
  // the first slice's kernel is exposed. Delta mode: one launch, no copies (see sf_block_host_delta). Work the caller has queued on OTHER streams for this handle must be complete (the Python
  // wrapper synchronises its stream first); everything queued here is complete on return.
  cudaStream_t sc = h->host_compute, sx = h->host_copy;
  // SF_FLAG_HOST_DELTA: h_obs is page-locked and still holds what the previous call with this flag wrote there
  uint4* host_alias = nullptr;
  sf_handle::Mirror* mir = nullptr;
  const size_t obs_bytes = n * per;
  const bool want_delta = render && (flags & SF_FLAG_HOST_DELTA);
  if (want_delta) {
    host_alias = reinterpret_cast<uint4*>(device_alias(h_obs));
    if (!host_alias) return fail(SF_ERR_INVALID, "SF_FLAG_HOST_DELTA needs a page-locked h_obs (sf_host_alloc, cudaHostAlloc, cudaHostRegister)");
    if (((uintptr_t)h_obs | (uintptr_t)host_alias) & 15) return fail(SF_ERR_INVALID, "SF_FLAG_HOST_DELTA needs a 16-byte aligned h_obs");
    for (auto& m : h->mirrors) if (m.host == h_obs) mir = &m;
    if (!mir) {  // a buffer seen for the first time: a free entry, else the least recently used one
      for (auto& m : h->mirrors) if (!mir || (mir->host && (!m.host || m.used < mir->used))) mir = &m;
      mir->host = nullptr;
    }
    if (obs_bytes > mir->cap) {
      if (mir->d) { cudaFree(mir->d); mir->d = nullptr; mir->cap = 0; }
      mir->host = nullptr;
      CUDA_TRY(cudaMalloc(&mir->d, obs_bytes));
      mir->cap = obs_bytes;
    }
    mir->used = ++h->mirror_clock;
    if (!h->d_delta_stats) {
      CUDA_TRY(cudaMalloc(&h->d_delta_stats, 16));
      CUDA_TRY(cudaMemset(h->d_delta_stats, 0, 16));
    }
  }
  const bool delta = want_delta && mir->host == h_obs && mir->bytes == obs_bytes;
  flags &= ~SF_FLAG_HOST_DELTA;
  SF_HP(1);
  CUDA_TRY(cudaStreamSynchronize(cudaStreamLegacy));  // calls made with stream == NULL (sf_reset, sf_seed ...) come first
  SF_HP(2);
  // page-locked actions / rewards / dones / kills / events are read and written in place by the kernel
  const int* k_actions = (const int*)device_alias(h_actions);
  int* k_reward = (int*)device_alias(h_reward);
  unsigned char* k_done = (unsigned char*)device_alias(h_done);
  unsigned char* k_kill = (unsigned char*)device_alias(h_fortkill);
  unsigned* k_events = (unsigned*)device_alias(h_events);
  SF_HP(3);
  if (!k_actions) CUDA_TRY(cudaMemcpyAsync(h->d_actions, h_actions, n * 4, cudaMemcpyHostToDevice, sc));
  const int slices = (render && !delta && n >= 2048) ? h->host_slices : 1;
  const bool fused_delta = delta && per % 16 == 0;  // the blocks of the step kernel send their own frames' changes
  for (int k = 0; k < slices; k++) {
    const size_t e0 = n * k / slices, e1 = n * (k + 1) / slices;
    SfRollArgs a = {};
    a.T = 1; a.EB = 1; a.ngroups = 0; a.flags = render ? flags : (flags & ~SF_FLAG_RENDER); a.action_seed = 0; a.t0 = 0;
    a.actions = k_actions ? k_actions : h->d_actions;
    a.obs = render ? h->d_obs : nullptr; a.reward = k_reward ? k_reward : h->d_reward; a.done = k_done ? k_done : h->d_done;
    a.fortkill = k_kill ? k_kill : h->d_kill; a.events = k_events ? k_events : h->d_events; a.sched = nullptr;
    a.env0 = (int)e0; a.envn = (int)(e1 - e0);
    if (fused_delta) { a.host_obs = host_alias; a.mirror = reinterpret_cast<uint4*>(mir->d); a.delta_stats = h->d_delta_stats; a.delta_lanes = h->delta_lanes; }
    rc = launch_rollout(h, a, sc);
    if (rc) { cudaStreamSynchronize(sc); cudaStreamSynchronize(sx); return rc; }
    if (render && !delta) {
      CUDA_TRY(cudaEventRecord(h->host_ev[k], sc));
      CUDA_TRY(cudaStreamWaitEvent(sx, h->host_ev[k], 0));
      CUDA_TRY(cudaMemcpyAsync(h_obs + e0 * per, h->d_obs + e0 * per, (e1 - e0) * per, cudaMemcpyDeviceToHost, sx));
    }
  }
  if (delta && !fused_delta) {
    const size_t n16 = obs_bytes / 16;
    const int blocks = (int)std::min<size_t>((n16 + 255) / 256, (size_t)h->num_sms * 8);
    auto kern = h->delta_lanes == 1 ? sf_host_delta_kernel<1> : (h->delta_lanes >= 4 ? sf_host_delta_kernel<4> : sf_host_delta_kernel<2>);
    kern<<<std::max(blocks, 1), 256, 0, sc>>>(reinterpret_cast<const uint4*>(h->d_obs), reinterpret_cast<uint4*>(mir->d), host_alias,
                                              n16, (int)(obs_bytes & 15), h->d_delta_stats);
    CUDA_TRY(cudaGetLastError());
  } else if (want_delta) {  // whole frames went out: from here on the mirror describes this buffer
    CUDA_TRY(cudaMemcpyAsync(mir->d, h->d_obs, obs_bytes, cudaMemcpyDeviceToDevice, sc));
    mir->host = h_obs; mir->bytes = obs_bytes;
  }
  if (render && !delta) h->full_calls++;
  if (h_reward && !k_reward) CUDA_TRY(cudaMemcpyAsync(h_reward, h->d_reward, n * 4, cudaMemcpyDeviceToHost, sc));
  if (h_done && !k_done) CUDA_TRY(cudaMemcpyAsync(h_done, h->d_done, n, cudaMemcpyDeviceToHost, sc));
  if (h_fortkill && !k_kill) CUDA_TRY(cudaMemcpyAsync(h_fortkill, h->d_kill, n, cudaMemcpyDeviceToHost, sc));
  if (h_events && !k_events) CUDA_TRY(cudaMemcpyAsync(h_events, h->d_events, n * 4, cudaMemcpyDeviceToHost, sc));
  SF_HP(4);
  cudaError_t e1 = cudaStreamSynchronize(sc), e2 = (render && !delta) ? cudaStreamSynchronize(sx) : cudaSuccess;
  SF_HP(5);
  if (e1 != cudaSuccess || e2 != cudaSuccess) {
    for (auto& m : h->mirrors) m.host = nullptr;
    return fail(SF_ERR_CUDA, std::string("sf_step_host: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2));
  }
  return SF_OK;
}

extern "C" int sf_host_forget(sf_handle* h, const void* h_obs) {
  if (!h) return fail(SF_ERR_INVALID, "handle is NULL");
  for (auto& m : h->mirrors) if (m.host == h_obs || !h_obs) m.host = nullptr;  // the device memory is kept for the next buffer
  return SF_OK;
}

extern "C" int sf_host_delta_stats(sf_handle* h, unsigned long long* out3) {
  if (!h || !out3) return fail(SF_ERR_INVALID, "handle or out is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  out3[0] = out3[1] = 0; out3[2] = h->full_calls;
  if (h->d_delta_stats) {
    CUDA_TRY(cudaStreamSynchronize(h->host_compute));
    CUDA_TRY(cudaMemcpy(out3, h->d_delta_stats, 16, cudaMemcpyDeviceToHost));
  }
  return SF_OK;
}

extern "C" int sf_get_state(sf_handle* h, int first, int count, sf_state_record* h_out) {
  if (!h || !h_out || first < 0 || count < 0 || first + count > h->dev.n) return fail(SF_ERR_INVALID, "bad range");
  if (count == 0) return SF_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  sf_state_record* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(sf_state_record) * (size_t)count));
  cudaError_t ce = cudaDeviceSynchronize();  // every stream: the state may be in flight on the caller's stream
  if (ce == cudaSuccess) {
    sf_get_state_kernel<<<(count + 63) / 64, 64>>>(h->dev, first, count, d);
    ce = cudaMemcpy(h_out, d, sizeof(sf_state_record) * (size_t)count, cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  if (ce != cudaSuccess) return fail(SF_ERR_CUDA, std::string("get_state: ") + cudaGetErrorString(ce));
  return SF_OK;
}

extern "C" int sf_set_state(sf_handle* h, int first, int count, const sf_state_record* h_in) {
  if (!h || !h_in || first < 0 || count < 0 || first + count > h->dev.n) return fail(SF_ERR_INVALID, "bad range");
  for (int k = 0; k < count; k++) {
    const sf_state_record& r = h_in[k];
    if (r.ship_angle != (double)(int)r.ship_angle || r.ship_angle < 0 || r.ship_angle >= 360)
      return fail(SF_ERR_UNSUPPORTED, "ship_angle must be an integer degree in [0,360) (it always is in the reference: game.cpp:148,317-324)");
    double fa = r.fortress_angle / 10.0, fl = r.fortress_last_angle / 10.0;
    if (fa != (double)(int)fa || fa < 0 || fa >= 36 || fl != (double)(int)fl || fl < 0 || fl >= 36)
      return fail(SF_ERR_UNSUPPORTED, "fortress angles must be multiples of 10 in [0,360) (game.cpp:40-41,206)");
    if (r.shell_mask >> SF_DEV_SHELLS) return fail(SF_ERR_UNSUPPORTED, "shell slots >= 4 can never be alive (at most 3 shells are in flight)");
    if (r.missile_mask >> SF_MAX_MISSILES) return fail(SF_ERR_INVALID, "missile_mask has more than 20 bits");
    for (int s = 0; s < SF_MAX_MISSILES; s++) if ((r.missile_mask >> s) & 1) {
      double a = r.missile_angle[s];
      if (a != (double)(int)a || a < 0 || a >= 360) return fail(SF_ERR_UNSUPPORTED, "missile angles must be integer degrees in [0,360)");
    }
  }
  if (count == 0) return SF_OK;
  CUDA_TRY(cudaSetDevice(h->device));
  sf_state_record* d = nullptr;
  CUDA_TRY(cudaMalloc(&d, sizeof(sf_state_record) * (size_t)count));
  cudaError_t ce = cudaMemcpy(d, h_in, sizeof(sf_state_record) * (size_t)count, cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
  if (ce == cudaSuccess) {
    sf_set_state_kernel<<<(count + 63) / 64, 64>>>(h->dev, first, count, d);
    ce = cudaDeviceSynchronize();
  }
  cudaFree(d);
  if (ce != cudaSuccess) return fail(SF_ERR_CUDA, std::string("set_state: ") + cudaGetErrorString(ce));
  return SF_OK;
}

__global__ void sf_epi_copy_kernel(unsigned long long* epi, long long* out, int reset) {
  int k = threadIdx.x;
  if (k < SF_NUM_EPISODE_STATS) {
    out[k] = (long long)epi[k];
    if (reset) epi[k] = 0;
  }
}

extern "C" int sf_episode_stats(sf_handle* h, long long* d_out, int reset, void* stream) {
  if (!h || !d_out) return fail(SF_ERR_INVALID, "handle or out is NULL");
  CUDA_TRY(cudaSetDevice(h->device));
  sf_epi_copy_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(h->dev.epi, d_out, reset);
  CUDA_TRY(cudaGetLastError());
  return SF_OK;
}

#ifdef SF_BARRIER_TIMING
extern "C" int sf_barrier_cycles(unsigned long long* h_out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h_out, sf_bar_cycles, sizeof(unsigned long long) * 16);
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(sf_bar_cycles, z, sizeof(z)); }
  return SF_OK;
}
#endif
#ifdef SF_TIMELINE
extern "C" int sf_timeline(unsigned long long* h_out) {  // [160][2][SF_TL_MAX]: (globaltimer ns << 8) | tag; zeroed after the read
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h_out, sf_tl, sizeof(unsigned long long) * 160 * 2 * SF_TL_MAX);
  static unsigned long long z[160 * 2 * SF_TL_MAX];
  cudaMemcpyToSymbol(sf_tl, z, sizeof(z));
  return SF_TL_MAX;
}
#endif
#ifdef SF_PHASE_TIMING
extern "C" int sf_debug_cycles(unsigned long long* h_out, int reset) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(h_out, sf_dbg_cycles, sizeof(unsigned long long) * 96);
  if (reset) { unsigned long long z[96] = {0}; cudaMemcpyToSymbol(sf_dbg_cycles, z, sizeof(z)); }
  return SF_OK;
}
#endif

extern "C" int sf_set_glyph_masks(sf_handle* h, const uint8_t* h_alpha, const uint8_t* h_slot) {
  if (!h || (!h_alpha) != (!h_slot)) return fail(SF_ERR_INVALID, "handle is NULL, or only one of alpha / slot given");
  SfTables* nt = new SfTables();
  char err[256] = {0};
  if (sf_build_tables_glyphs(nt, err, sizeof(err), h_alpha, h_slot)) { std::string m = err; delete nt; return fail(SF_ERR_INVALID, "table build failed: " + m); }
  cudaError_t ce = cudaSetDevice(h->device);
  if (ce == cudaSuccess) ce = cudaDeviceSynchronize();  // nothing may be drawing from the old tables
  if (ce == cudaSuccess) ce = cudaMemcpy((void*)h->dev.tab, nt, sizeof(SfTables), cudaMemcpyHostToDevice);
  if (ce == cudaSuccess) ce = cudaMemset(h->dev.expo_meta, 0, sizeof(uint2) * (size_t)h->dev.n_pad);  // cached resampled boxes may hold old digits
  if (ce == cudaSuccess) {
    sf_pack_static_kernel<<<1, SF_BLOCK, SF_RENDER_SMEM_BYTES(SF_WARPS_PER_BLOCK)>>>(h->dev, const_cast<unsigned char*>(h->dev.static_image));
    ce = cudaGetLastError();
    if (ce == cudaSuccess) ce = cudaDeviceSynchronize();
  }
  if (ce != cudaSuccess) { delete nt; return fail(SF_ERR_CUDA, std::string("sf_set_glyph_masks: ") + cudaGetErrorString(ce)); }
  delete h->h_tab;
  h->h_tab = nt;
  return SF_OK;
}

extern "C" int sf_background(const sf_handle* h, uint8_t* h_native, uint8_t* h_obs) {
  if (!h) return fail(SF_ERR_INVALID, "handle is NULL");
  if (h_native)
    for (int r = 0; r < SF_NAT_H; r++) memcpy(h_native + r * SF_NAT_W, h->h_tab->bg_nat + r * SF_NAT_STRIDE, SF_NAT_W);
  if (h_obs) memcpy(h_obs, h->h_tab->bg_obs, 84 * 84);
  return SF_OK;
}
