// tt_replay.cu -- kernel (d): vectorised ring-buffer scatter of (s, a, r, s', done) and the sample gather.
// Replaces ReplayBuffer.store_transition / sample_buffer (DDPG/replay_buffer.py:13-34) for n transitions.
//
// Batched store == n sequential store_transition calls in env order: transition i goes to ring row
// (mem_cntr + i) % mem_size and, when n > mem_size, the LAST writer of a row wins; so only the last
// min(n, mem_size) transitions are written, each exactly once -> no write races.
// The ring arrays are dense float32 [mem_size, 23]; a run of consecutive rows is a flat float range, so the
// copy is done element-wise over 23*rows floats with unit-stride (coalesced) loads and stores, using 16 B
// vectors when source and destination of a segment are both 16 B aligned.
#include "tt_common.cuh"

namespace {
constexpr int kThreads = 256;

__device__ __forceinline__ void copy_rows(float *__restrict__ dst, const float *__restrict__ src, int64_t nfloat,
                                          int64_t tid, int64_t nth) {
    if (((reinterpret_cast<uintptr_t>(dst) | reinterpret_cast<uintptr_t>(src)) & 15) == 0) {
        const int64_t nv = nfloat >> 2;
        const float4 *s4 = reinterpret_cast<const float4 *>(src);
        float4 *d4 = reinterpret_cast<float4 *>(dst);
        for (int64_t v = tid; v < nv; v += nth) __stcs(&d4[v], __ldcs(&s4[v]));
        for (int64_t v = (nv << 2) + tid; v < nfloat; v += nth) dst[v] = src[v];
    } else {
        for (int64_t v = tid; v < nfloat; v += nth) __stcs(&dst[v], __ldcs(&src[v]));
    }
}

// segment = transitions [i0, i0 + cnt) -> ring rows [r0, r0 + cnt) (contiguous, no wrap inside a segment)
struct Seg { int64_t i0, r0, cnt; };

__global__ void __launch_bounds__(kThreads) replay_store_kernel(float *__restrict__ S, float *__restrict__ A, float *__restrict__ R,
                                                                float *__restrict__ S2, uint8_t *__restrict__ D,
                                                                const float *__restrict__ s, int64_t ld_s,
                                                                const float *__restrict__ a, const float *__restrict__ r,
                                                                const float *__restrict__ s2, int64_t ld_s2,
                                                                const uint8_t *__restrict__ d, Seg g0, Seg g1) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const Seg g = k ? g1 : g0;
        if (g.cnt <= 0) continue;
        if (ld_s == TT_OBS_DIM) copy_rows(S + g.r0 * TT_OBS_DIM, s + g.i0 * TT_OBS_DIM, g.cnt * TT_OBS_DIM, tid, nth);
        else for (int64_t v = tid; v < g.cnt * TT_OBS_DIM; v += nth) {
            const int64_t rr = v / TT_OBS_DIM, c = v - rr * TT_OBS_DIM;
            S[(g.r0 + rr) * TT_OBS_DIM + c] = s[(g.i0 + rr) * ld_s + c];
        }
        if (ld_s2 == TT_OBS_DIM) copy_rows(S2 + g.r0 * TT_OBS_DIM, s2 + g.i0 * TT_OBS_DIM, g.cnt * TT_OBS_DIM, tid, nth);
        else for (int64_t v = tid; v < g.cnt * TT_OBS_DIM; v += nth) {
            const int64_t rr = v / TT_OBS_DIM, c = v - rr * TT_OBS_DIM;
            S2[(g.r0 + rr) * TT_OBS_DIM + c] = s2[(g.i0 + rr) * ld_s2 + c];
        }
        copy_rows(A + g.r0, a + g.i0, g.cnt, tid, nth);
        copy_rows(R + g.r0, r + g.i0, g.cnt, tid, nth);
        for (int64_t v = tid; v < g.cnt; v += nth) D[g.r0 + v] = d[g.i0 + v];
    }
}

__global__ void __launch_bounds__(kThreads) replay_gather_kernel(const float *__restrict__ S, const float *__restrict__ A,
                                                                 const float *__restrict__ R, const float *__restrict__ S2,
                                                                 const uint8_t *__restrict__ D, const int64_t *__restrict__ rows,
                                                                 int64_t batch, float *__restrict__ s, float *__restrict__ a,
                                                                 float *__restrict__ r, float *__restrict__ s2,
                                                                 uint8_t *__restrict__ d) {
    const int64_t tid = (int64_t)blockIdx.x * kThreads + threadIdx.x, nth = (int64_t)gridDim.x * kThreads;
    for (int64_t v = tid; v < batch * TT_OBS_DIM; v += nth) {
        const int64_t b = v / TT_OBS_DIM, c = v - b * TT_OBS_DIM, row = rows[b];
        s[v] = S[row * TT_OBS_DIM + c];
        s2[v] = S2[row * TT_OBS_DIM + c];
    }
    for (int64_t b = tid; b < batch; b += nth) {
        const int64_t row = rows[b];
        a[b] = A[row]; r[b] = R[row]; d[b] = D[row];
    }
}
}  // namespace

namespace tt {
int replay_store(float *S, float *A, float *R, float *S2, uint8_t *D, int64_t cap, int64_t cntr, const float *s, int64_t ld_s,
                 const float *a, const float *r, const float *s2, int64_t ld_s2, const uint8_t *d, int64_t n, cudaStream_t st) {
    // only the last min(n, cap) transitions survive; they occupy at most two contiguous row ranges
    const int64_t first = n > cap ? n - cap : 0, live = n - first;
    const int64_t r0 = (cntr + first) % cap;
    Seg g0, g1;
    g0.i0 = first; g0.r0 = r0; g0.cnt = live < cap - r0 ? live : cap - r0;
    g1.i0 = first + g0.cnt; g1.r0 = 0; g1.cnt = live - g0.cnt;
    const int64_t work = live * TT_OBS_DIM / 4 + 1;
    int64_t blocks = (work + kThreads - 1) / kThreads;
    const int64_t cap_blocks = (int64_t)tt::sm_count() * 16;
    if (blocks > cap_blocks) blocks = cap_blocks;
    replay_store_kernel<<<(unsigned)blocks, kThreads, 0, st>>>(S, A, R, S2, D, s, ld_s, a, r, s2, ld_s2, d, g0, g1);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}
}  // namespace tt

extern "C" {

int tt_replay_store(float *d_state_mem, float *d_action_mem, float *d_reward_mem, float *d_new_state_mem,
                    uint8_t *d_terminal_mem, int64_t mem_size, int64_t mem_cntr, const float *d_s, int64_t ld_s,
                    const float *d_a, const float *d_r, const float *d_s2, int64_t ld_s2, const uint8_t *d_done, int64_t n,
                    tt_stream_t stream) {
    TT_REQUIRE(d_state_mem && d_action_mem && d_reward_mem && d_new_state_mem && d_terminal_mem, "NULL ring pointer");
    TT_REQUIRE(d_s && d_a && d_r && d_s2 && d_done, "NULL transition pointer");
    TT_REQUIRE(mem_size > 0 && mem_cntr >= 0 && n > 0, "bad sizes");
    TT_REQUIRE(ld_s >= TT_OBS_DIM && ld_s2 >= TT_OBS_DIM, "ld < 23");
    return tt::replay_store(d_state_mem, d_action_mem, d_reward_mem, d_new_state_mem, d_terminal_mem, mem_size, mem_cntr, d_s,
                            ld_s, d_a, d_r, d_s2, ld_s2, d_done, n, tt::as_stream(stream));
}

int tt_replay_gather(const float *d_state_mem, const float *d_action_mem, const float *d_reward_mem,
                     const float *d_new_state_mem, const uint8_t *d_terminal_mem, const int64_t *d_rows, int64_t batch,
                     float *d_s, float *d_a, float *d_r, float *d_s2, uint8_t *d_done, tt_stream_t stream) {
    TT_REQUIRE(d_state_mem && d_action_mem && d_reward_mem && d_new_state_mem && d_terminal_mem && d_rows, "NULL argument");
    TT_REQUIRE(d_s && d_a && d_r && d_s2 && d_done && batch > 0, "NULL output / bad batch");
    int64_t blocks = (batch * TT_OBS_DIM + kThreads - 1) / kThreads;
    if (blocks > 148 * 8) blocks = 148 * 8;
    replay_gather_kernel<<<(unsigned)blocks, kThreads, 0, tt::as_stream(stream)>>>(d_state_mem, d_action_mem, d_reward_mem,
                                                                                  d_new_state_mem, d_terminal_mem, d_rows, batch,
                                                                                  d_s, d_a, d_r, d_s2, d_done);
    TT_COUNT_LAUNCH(); TT_LAUNCH_CHECK();
    return TT_OK;
}

}  // extern "C"
