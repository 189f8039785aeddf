// Device-resident vectorised environments: one kernel per vector step replaces the reference's host
// `envs.step` + masked `envs.reset` + list append (diamond/ppo.py:160-182) and both PCIe crossings of a rollout step.
// The kernel steps every environment, writes the step straight into row t of the rollout buffer with the casts of
// ppo.py:229-232, keeps Gymnasium's autoreset-DISABLED contract (next_obs is the TRUE final observation; environments
// that finished are then reset, and the post-reset observation is what the next sampling step sees, ppo.py:174-179), and
// maintains per-environment episode return / length.
//
// Dynamics follow Gymnasium's classic-control environments (third-party, not in the reference tree; gymnasium >= 1.0.0,
// `classic_control/cartpole.py` CartPole-v1: Euler integrator, tau 0.02, termination |x| > 2.4 or |theta| > 12 deg, 500-step
// time limit; `classic_control/pendulum.py` Pendulum-v1: dt 0.05, g 10, torque clipped to [-2, 2], speed clipped to [-8, 8],
// 200-step time limit, never terminates) with the state held in float64 as Gymnasium does, observations cast to float32.
// The checker is oracle/env_oracle.py; reset draws come from Philox4x32-10 keyed by (seed, global env id, episode index),
// so a run does not depend on launch geometry or GPU count.
#include "common.cuh"

namespace {

__device__ __forceinline__ uint4 philox4x32_10(uint4 ctr, uint2 key)
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
__device__ __forceinline__ double u01d(uint32_t x) { return ((double)x + 0.5) * (1.0 / 4294967296.0); }      // (0,1)

struct EnvArgs {
    int kind, N, D, A, t, auto_reset;
    double* state;              // [N, 4]  CartPole: x, x_dot, theta, theta_dot; Pendulum: theta, theta_dot; synthetic: unused
    int32_t* steps;             // [N] steps taken in the current episode
    int64_t* episode;           // [N] episodes started so far (reset draw index)
    double* ep_return;          // [N] running return
    const void* actions;        // int64 [N] (discrete) or f32 [N, A]
    float *obs_t, *next_obs_t;  // row t of the rollout buffer: [N, D] each (obs_t may be null)
    void* act_t;                // row t: int32 [N] or f32 [N, A] (may be null)
    float *rew_t, *term_t, *trunc_t;
    float* cur_obs;             // [N, D] observation the next sampling step reads
    float* done_return;         // [N] return of the episode that ended at this step (else untouched), optional
    unsigned long long seed;
    long long env_offset;
    float p_term, p_trunc;      // synthetic env
};

__device__ __forceinline__ void reset_state(const EnvArgs& a, int e, double* s)
{
    const unsigned long long env = (unsigned long long)(a.env_offset + e);
    const unsigned long long ep = (unsigned long long)a.episode[e];
    const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)ep, (uint32_t)(ep >> 32) ^ 0x52534554u),
                                  make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
    if (a.kind == 0) {            // uniform(-0.05, 0.05) x 4
        s[0] = -0.05 + 0.1 * u01d(r.x); s[1] = -0.05 + 0.1 * u01d(r.y); s[2] = -0.05 + 0.1 * u01d(r.z); s[3] = -0.05 + 0.1 * u01d(r.w);
    } else {                      // theta ~ U(-pi, pi), theta_dot ~ U(-1, 1)
        s[0] = -3.14159265358979323846 + 6.28318530717958647692 * u01d(r.x); s[1] = -1.0 + 2.0 * u01d(r.y); s[2] = 0.0; s[3] = 0.0;
    }
    a.episode[e] += 1;
    a.steps[e] = 0;
}

__device__ __forceinline__ void write_obs(int kind, const double* s, float* o)
{
    if (kind == 0) { o[0] = (float)s[0]; o[1] = (float)s[1]; o[2] = (float)s[2]; o[3] = (float)s[3]; }
    else { o[0] = (float)cos(s[0]); o[1] = (float)sin(s[0]); o[2] = (float)s[1]; }
}

// mode 0: step (t, actions); mode 1: reset the environments with mask[e] != 0 (mask null: all)
__global__ void classic_env_kernel(EnvArgs a, int mode, const unsigned char* __restrict__ mask)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= a.N) return;
    double s[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) s[i] = a.state[(int64_t)e * 4 + i];
    if (mode == 1) {
        if (mask == nullptr || mask[e]) {
            reset_state(a, e, s);
            a.ep_return[e] = 0.0;
#pragma unroll
            for (int i = 0; i < 4; ++i) a.state[(int64_t)e * 4 + i] = s[i];
            write_obs(a.kind, s, a.cur_obs + (int64_t)e * a.D);
        }
        return;
    }
    if (a.obs_t) write_obs(a.kind, s, a.obs_t + (int64_t)e * a.D);
    double reward;
    bool term = false, trunc;
    if (a.kind == 0) {
        const long long act = reinterpret_cast<const long long*>(a.actions)[e];
        if (a.act_t) reinterpret_cast<int32_t*>(a.act_t)[e] = (int32_t)act;
        const double gravity = 9.8, masscart = 1.0, masspole = 0.1, length = 0.5, force_mag = 10.0, tau = 0.02;
        const double total_mass = masspole + masscart, pml = masspole * length;
        const double force = act == 1 ? force_mag : -force_mag;
        const double ct = cos(s[2]), st = sin(s[2]);
        const double temp = (force + pml * s[3] * s[3] * st) / total_mass;
        const double th_acc = (gravity * st - ct * temp) / (length * (4.0 / 3.0 - masspole * ct * ct / total_mass));
        const double x_acc = temp - pml * th_acc * ct / total_mass;
        const double x = s[0] + tau * s[1], xd = s[1] + tau * x_acc, th = s[2] + tau * s[3], thd = s[3] + tau * th_acc;
        s[0] = x; s[1] = xd; s[2] = th; s[3] = thd;
        term = fabs(x) > 2.4 || fabs(th) > 12.0 * 2.0 * 3.14159265358979323846 / 360.0;
        reward = 1.0;
        a.steps[e] += 1;
        trunc = a.steps[e] >= 500 && !term;
    } else {
        const float af = reinterpret_cast<const float*>(a.actions)[(int64_t)e * a.A];
        if (a.act_t) reinterpret_cast<float*>(a.act_t)[(int64_t)e * a.A] = af;
        const double max_speed = 8.0, max_torque = 2.0, dt = 0.05, g = 10.0, m = 1.0, l = 1.0, pi = 3.14159265358979323846;
        const double u = fmin(fmax((double)af, -max_torque), max_torque);
        double thn = fmod(s[0] + pi, 2.0 * pi);
        if (thn < 0.0) thn += 2.0 * pi;                                 // python's % is non-negative
        thn -= pi;
        reward = -(thn * thn + 0.1 * s[1] * s[1] + 0.001 * u * u);
        double thd = s[1] + (3.0 * g / (2.0 * l) * sin(s[0]) + 3.0 / (m * l * l) * u) * dt;
        thd = fmin(fmax(thd, -max_speed), max_speed);
        s[1] = thd;
        s[0] = s[0] + thd * dt;
        a.steps[e] += 1;
        trunc = a.steps[e] >= 200;
    }
    write_obs(a.kind, s, a.next_obs_t + (int64_t)e * a.D);             // the true final observation (autoreset disabled)
    a.rew_t[e] = (float)reward;
    a.term_t[e] = term ? 1.0f : 0.0f;
    a.trunc_t[e] = trunc ? 1.0f : 0.0f;
    const double ret = a.ep_return[e] + reward;
    const bool done = term || trunc;
    if (done && a.done_return) a.done_return[e] = (float)ret;
    if (done && a.auto_reset) {
        reset_state(a, e, s);                                           // envs.reset(options={"reset_mask": dones}), ppo.py:174-179
        a.ep_return[e] = 0.0;
    } else {
        a.ep_return[e] = ret;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) a.state[(int64_t)e * 4 + i] = s[i];
    write_obs(a.kind, s, a.cur_obs + (int64_t)e * a.D);
}

// Synthetic environment of the scale benchmark (shape-only stand-in, diamond/envs.py BatchedSyntheticVectorEnv): i.i.d.
// standard-normal observations, reward = 0.1 * obs[0] * sign(action), Bernoulli terminations / truncations.
// One warp per environment; lanes stride over the observation.
__global__ void synthetic_env_kernel(EnvArgs a, int mode, const unsigned char* __restrict__ mask)
{
    const int e = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
    if (e >= a.N) return;
    const unsigned long long env = (unsigned long long)(a.env_offset + e);
    auto fresh = [&](unsigned long long draw, float* dst0, float* dst1) {       // D normals into up to two destinations
        for (int j0 = 2 * lane; j0 < a.D; j0 += 64) {
            const unsigned long long c = draw * 4096ull + (unsigned long long)(j0 / 2);
            const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)c, (uint32_t)(c >> 32) ^ 0x4F425321u),
                                          make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
            const float rad = sqrtf(-2.0f * logf((float)u01d(r.x)));
            float sn, cs;
            sincosf(6.28318530717958647692f * (float)u01d(r.y), &sn, &cs);
            const float v0 = rad * cs, v1 = rad * sn;
            if (dst0) { dst0[j0] = v0; if (j0 + 1 < a.D) dst0[j0 + 1] = v1; }
            if (dst1) { dst1[j0] = v0; if (j0 + 1 < a.D) dst1[j0 + 1] = v1; }
        }
    };
    float* cur = a.cur_obs + (int64_t)e * a.D;
    if (mode == 1) {
        if (mask == nullptr || mask[e]) {
            const unsigned long long draw = 2ull * (unsigned long long)a.episode[e] + 1ull;
            fresh((draw << 20), cur, nullptr);
            __syncwarp();
            if (lane == 0) { a.episode[e] += 1; a.steps[e] = 0; a.ep_return[e] = 0.0; }
        }
        return;
    }
    const float o0 = cur[0];
    __syncwarp();
    if (a.obs_t) for (int j = lane; j < a.D; j += 32) a.obs_t[(int64_t)e * a.D + j] = cur[j];
    float asum;
    if (a.A > 0 && a.kind == 3) {                       // continuous synthetic: sum of the action vector
        asum = 0.f;
        for (int j = 0; j < a.A; ++j) {
            const float v = reinterpret_cast<const float*>(a.actions)[(int64_t)e * a.A + j];
            asum += v;
            if (a.act_t && lane == 0) reinterpret_cast<float*>(a.act_t)[(int64_t)e * a.A + j] = v;
        }
    } else {
        const long long act = reinterpret_cast<const long long*>(a.actions)[e];
        if (a.act_t && lane == 0) reinterpret_cast<int32_t*>(a.act_t)[e] = (int32_t)act;
        asum = (float)act;
    }
    const float reward = o0 * (asum > 0.f ? 1.0f : -1.0f) * 0.1f;
    const unsigned long long step_id = ((unsigned long long)a.episode[e] << 20) + (unsigned long long)a.steps[e];
    __syncwarp();
    // next observation: a fresh draw, written to the buffer's final-observation row and (if the episode continues) cur_obs
    const uint4 r = philox4x32_10(make_uint4((uint32_t)env, (uint32_t)(env >> 32), (uint32_t)step_id, (uint32_t)(step_id >> 32) ^ 0x444F4E45u),
                                  make_uint2((uint32_t)a.seed, (uint32_t)(a.seed >> 32)));
    const bool term = (float)u01d(r.x) < a.p_term;
    const bool trunc = !term && (float)u01d(r.y) < a.p_trunc;
    const bool done = term || trunc;
    fresh(2ull * step_id + 2ull, a.next_obs_t + (int64_t)e * a.D, (done && a.auto_reset) ? nullptr : cur);
    if (done && a.auto_reset) fresh(((2ull * (unsigned long long)(a.episode[e] + 1) + 1ull) << 20) + 7ull, cur, nullptr);
    __syncwarp();
    if (lane == 0) {
        a.rew_t[e] = reward;
        a.term_t[e] = term ? 1.0f : 0.0f;
        a.trunc_t[e] = trunc ? 1.0f : 0.0f;
        const double ret = a.ep_return[e] + (double)reward;
        if (done && a.done_return) a.done_return[e] = (float)ret;
        if (done && a.auto_reset) { a.episode[e] += 1; a.steps[e] = 0; a.ep_return[e] = 0.0; }
        else { a.steps[e] += 1; a.ep_return[e] = ret; }
    }
}

int launch_env(dppo_ctx* ctx, const EnvArgs& a, int mode, const unsigned char* mask, cudaStream_t st)
{
    if (a.kind == 0 || a.kind == 1) {
        classic_env_kernel<<<(a.N + 127) / 128, 128, 0, st>>>(a, mode, mask);
        DPPO_CHECK_LAUNCH(ctx, "classic_env_kernel");
    } else {
        synthetic_env_kernel<<<(unsigned)(((int64_t)a.N * 32 + 255) / 256), 256, 0, st>>>(a, mode, mask);
        DPPO_CHECK_LAUNCH(ctx, "synthetic_env_kernel");
    }
    return 0;
}

int fill_args(dppo_ctx* ctx, const dppo_env_desc* d, const dppo_env_state* s, EnvArgs& a)
{
    if (!d || !s) DPPO_FAIL(ctx, "env: null descriptor / state");
    if (d->kind < 0 || d->kind > 3 || d->num_envs < 1) DPPO_FAIL(ctx, "env: bad descriptor (kind %d, num_envs %d)", d->kind, d->num_envs);
    const int want_d = d->kind == 0 ? 4 : d->kind == 1 ? 3 : d->obs_dim;
    if (d->obs_dim != want_d || want_d < 1) DPPO_FAIL(ctx, "env: obs_dim %d does not match kind %d", d->obs_dim, d->kind);
    if (!s->state || !s->steps || !s->episode || !s->ep_return || !s->cur_obs) DPPO_FAIL(ctx, "env: null state buffer");
    a.kind = d->kind; a.N = d->num_envs; a.D = d->obs_dim; a.A = d->act_dim;
    a.state = s->state; a.steps = s->steps; a.episode = s->episode; a.ep_return = s->ep_return; a.cur_obs = s->cur_obs;
    a.seed = d->seed; a.env_offset = d->env_offset; a.p_term = d->p_term; a.p_trunc = d->p_trunc;
    return 0;
}

}  // namespace

extern "C" int dppo_env_reset(dppo_ctx* ctx, const dppo_env_desc* desc, const dppo_env_state* state, const unsigned char* mask,
                              void* stream)
{
    if (!ctx) return 1;
    EnvArgs a = {};
    if (fill_args(ctx, desc, state, a)) return 1;
    return launch_env(ctx, a, 1, mask, (cudaStream_t)stream);
}

extern "C" int dppo_env_step(dppo_ctx* ctx, const dppo_env_desc* desc, const dppo_env_state* state, const void* actions, int t,
                             int auto_reset, float* obs, float* next_obs, void* buf_actions, float* rewards, float* terminations,
                             float* truncations, float* done_return, void* stream)
{
    if (!ctx) return 1;
    EnvArgs a = {};
    if (fill_args(ctx, desc, state, a)) return 1;
    if (!actions || !next_obs || !rewards || !terminations || !truncations || t < 0) DPPO_FAIL(ctx, "env_step: null output / bad step");
    const int64_t N = a.N, D = a.D;
    const int cont = (a.kind == 1 || a.kind == 3);
    a.t = t; a.auto_reset = auto_reset; a.actions = actions;
    a.obs_t = obs ? obs + (int64_t)t * N * D : nullptr;
    a.next_obs_t = next_obs + (int64_t)t * N * D;
    a.act_t = buf_actions ? (cont ? (void*)((float*)buf_actions + (int64_t)t * N * a.A) : (void*)((int32_t*)buf_actions + (int64_t)t * N)) : nullptr;
    a.rew_t = rewards + (int64_t)t * N; a.term_t = terminations + (int64_t)t * N; a.trunc_t = truncations + (int64_t)t * N;
    a.done_return = done_return;
    return launch_env(ctx, a, 0, nullptr, (cudaStream_t)stream);
}
