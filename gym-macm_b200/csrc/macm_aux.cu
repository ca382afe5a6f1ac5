// macm_aux.cu -- kernels beside the step: the on-device initial-state sampler
// (Flock.__init__ / TDM.__init__ distributions, mvmnt.py:48-52,62-64; combat.py:84-86) and the
// scripted actors of test_scripts/bots.py (the `actions=None` mode of mvmnt.py:86-92).
#include "macm_sim.h"

namespace {

__global__ void macm_sample_kernel(const __grid_constant__ SimConst P, const __grid_constant__ SampleConst sc)
{
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t EN = (uint64_t)P.E * P.N, ET = (uint64_t)P.E * P.T;
    if (idx < EN) {
        float x, y, a;
        sample_agent(P, sc, idx + (uint64_t)P.env_base * P.N, (int)(idx % P.N), 0u, x, y, a);
        P.posvel[idx] = make_float4(x, y, 0.0f, 0.0f);
        P.angsleep[idx] = make_float2(a, 0.0f);
    } else if (idx < EN + ET) {
        const uint64_t t = idx - EN;
        const_cast<float2*>(P.targets)[t] = sample_target(sc, t + (uint64_t)P.env_base * P.T, 0u);
    }
}

// bots.py: idle/forward/rotate/diag (bots:19-29), circle (bots:31-35), flock (bots:37-61), combat (bots:3-16),
// plus U{0,1,2}^3 x U{0,1}
__global__ void macm_bot_kernel(const __grid_constant__ SimConst P, int policy, uint64_t seed, uint32_t* out)
{
    const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= (uint64_t)P.E * P.N) return;
    uint32_t a0 = 1, a1 = 1, a2 = 1, a3 = 0;
    switch (policy) {
        case MACM_BOT_FORWARD: a0 = 2; break;
        case MACM_BOT_ROTATE: a2 = 2; break;
        case MACM_BOT_DIAG: a0 = 2; a1 = 2; break;
        case MACM_BOT_RANDOM: {
            const int step = P.env_state[gi / P.N].x;
            const Philox r(seed, gi + (uint64_t)P.env_base * P.N, 2u, (uint32_t)step);
            a0 = (uint32_t)(((uint64_t)r.c[0] * 3u) >> 32);
            a1 = (uint32_t)(((uint64_t)r.c[1] * 3u) >> 32);
            a2 = (uint32_t)(((uint64_t)r.c[2] * 3u) >> 32);
            a3 = P.kind == MACM_ENV_TDM ? (r.c[3] >> 31) : 0u;
            break;
        }
        case MACM_BOT_FLOCK: {
            // steer towards the target node: position = [r, theta] or [r, cos, sin]
            float r, th_sign, ahead;
            if (P.coord == MACM_COORD_POLAR) {
                const float4 o = reinterpret_cast<const float4*>(P.obs)[gi];
                r = o.z;
                th_sign = (o.w > 0.0f) ? 1.0f : (o.w < 0.0f ? -1.0f : 0.0f);          // np.sign(theta)
                ahead = fabsf(o.w) < (float)(NP_PI / 4) ? 1.0f : 0.0f;                 // |theta| < pi/4
            } else {
                const float* o = P.obs + gi * 6;
                r = o[3];
                th_sign = (o[5] > 0.0f) ? 1.0f : (o[5] < 0.0f ? -1.0f : 0.0f);        // np.sign(sin)
                ahead = o[4] > 0.70710678f ? 1.0f : 0.0f;                              // cos > cos(pi/4)
            }
            if (!(r < 1.0f)) { a2 = (uint32_t)(th_sign + 1.0f); a0 = (uint32_t)(ahead + 1.0f); }
            break;
        }
        case MACM_BOT_CIRCLE: {   // bots.py:31-35: forward, turning on a coin flip
            const int step = P.env_state[gi / P.N].x;
            const Philox r(seed, gi + (uint64_t)P.env_base * P.N, 3u, (uint32_t)step);
            a0 = 2; a2 = (r.c[0] >> 31) ? 2u : 1u;
            break;
        }
        case MACM_BOT_COMBAT: {
            // bots.py:3-16 on the agent's observation row (combat.py:206-227): the nearest enemy (strict '<' in
            // list order: the lowest index among equals), turn towards it, walk when it is within +-36 degrees,
            // strike inside 3 m.  A dead agent's row has no entries: idle.
            const int N = P.N;
            const float4* row = reinterpret_cast<const float4*>(P.obs) + gi * N;
            float br = 0.0f, bth = 0.0f;
            bool found = false;
            for (int j = 0; j < N; ++j) {
                const float4 o = row[j];
                if (o.w == 0.0f && (!found || o.x < br)) { br = o.x; bth = o.y; found = true; }
            }
            if (found) {
                a0 = (fabs((double)bth) < NP_PI / 5) ? 2u : 1u;
                a2 = bth > 0.0f ? 2u : (bth < 0.0f ? 0u : 1u);
                a3 = br < 3.0f ? 1u : 0u;
            }
            break;
        }
        default: break;
    }
    out[gi] = a0 | (a1 << 8) | (a2 << 16) | (a3 << 24);
}

// env_state[:,1] overflow bits -> two counters
__global__ void macm_overflow_kernel(const __grid_constant__ SimConst P, int* out2)
{
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = e < P.E ? P.env_state[e].y : 0;
    const unsigned c = __ballot_sync(0xffffffffu, (f & MACM_ENV_CONTACT_OVERFLOW) != 0);
    const unsigned t = __ballot_sync(0xffffffffu, (f & MACM_ENV_TOUCH_OVERFLOW) != 0);
    if ((threadIdx.x & 31) == 0) {
        if (c) atomicAdd(&out2[0], __popc(c));
        if (t) atomicAdd(&out2[1], __popc(t));
    }
}

// any integer tensor [.., width] (width 3 or 4; elements of 1, 2, 4 or 8 bytes, little endian) -> the uint8 [.., 4]
// action words of macm_step
__global__ void macm_pack_kernel(const unsigned char* __restrict__ src, int elem_bytes, int width, uint64_t n, uint32_t* out)
{
    const uint64_t gi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gi >= n) return;
    uint32_t w = 0u;
    for (int k = 0; k < width; ++k) w |= (uint32_t)src[(gi * width + k) * elem_bytes] << (8 * k);
    out[gi] = w;
}

}  // namespace

cudaError_t macm_launch_sample(const SimConst& P, const SampleConst& sc, cudaStream_t s)
{
    const uint64_t total = (uint64_t)P.E * P.N + (uint64_t)P.E * P.T;
    const int threads = 256;
    macm_sample_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(P, sc);
    return cudaGetLastError();
}

cudaError_t macm_launch_overflow_count(const SimConst& P, int* d_out2, cudaStream_t s)
{
    cudaError_t e = cudaMemsetAsync(d_out2, 0, 2 * sizeof(int), s);
    if (e != cudaSuccess) return e;
    macm_overflow_kernel<<<(unsigned)((P.E + 255) / 256), 256, 0, s>>>(P, d_out2);
    return cudaGetLastError();
}

cudaError_t macm_launch_pack_actions(const void* src, int elem_bytes, int width, uint64_t n_agents, void* out, cudaStream_t s)
{
    macm_pack_kernel<<<(unsigned)((n_agents + 255) / 256), 256, 0, s>>>((const unsigned char*)src, elem_bytes, width,
                                                                          n_agents, (uint32_t*)out);
    return cudaGetLastError();
}

cudaError_t macm_launch_bot(const SimConst& P, int policy, uint64_t seed, void* actions_out, cudaStream_t s)
{
    const uint64_t total = (uint64_t)P.E * P.N;
    const int threads = 256;
    macm_bot_kernel<<<(unsigned)((total + threads - 1) / threads), threads, 0, s>>>(P, policy, seed,
                                                                                    (uint32_t*)actions_out);
    return cudaGetLastError();
}
