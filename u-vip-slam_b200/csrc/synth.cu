// synth.cu — SURVEY.md Appendix B `synth_frame` on the device (bench / test utility of libuvip_orb.so, not part of the hot path).
// BASELINE config 5 is 8 sequences x 4096 frames of 1280x1024 (43 GB of input): the host generator (u-vip-slam_b200/synth.py, numpy)
// takes ~0.15 s per frame, this kernel ~0.3 ms, with identical bytes (integer-only, counter-based SplitMix64; checked against the
// numpy generator in tests/test_gpu_parity.py).
#include "common.cuh"

namespace uvip {

__host__ __device__ __forceinline__ unsigned long long splitmix_draw(unsigned long long seed, unsigned long long k)
{
    unsigned long long z = seed + k * 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

constexpr int SYN_MAXR = 2048;            // rectangles per frame: (W + 64)(H + 64) / 1500

__global__ void __launch_bounds__(256)
k_synth_frames(const long long* __restrict__ seeds, const int* __restrict__ dxy, const long long* __restrict__ noise_seeds, int W, int H,
               uint8_t* __restrict__ out, size_t frame_pitch)
{
    __shared__ unsigned short s_x0[SYN_MAXR], s_y0[SYN_MAXR], s_x1[SYN_MAXR], s_y1[SYN_MAXR];
    __shared__ unsigned char s_g[SYN_MAXR];
    const int f = blockIdx.y;
    const unsigned long long seed = (unsigned long long)seeds[f], nseed = (unsigned long long)noise_seeds[f] ^ 0xA5A5A5A5ULL;
    const int dx = dxy[2 * f], dy = dxy[2 * f + 1];
    const int Wc = W + 64, Hc = H + 64, gh = Hc / 16 + 2, gw = Wc / 16 + 2;
    const int R = (int)(((long long)Wc * Hc) / 1500);
    const unsigned long long base = (unsigned long long)gh * gw + 1ULL;
    for (int i = threadIdx.x; i < R; i += 256) {
        const unsigned long long b = base + 5ULL * i;
        const int rx = (int)(splitmix_draw(seed, b) % (unsigned long long)Wc), ry = (int)(splitmix_draw(seed, b + 1) % (unsigned long long)Hc);
        const int rw = 6 + (int)(splitmix_draw(seed, b + 2) % 43ULL), rh = 6 + (int)(splitmix_draw(seed, b + 3) % 43ULL);
        s_x0[i] = (unsigned short)rx; s_y0[i] = (unsigned short)ry;
        s_x1[i] = (unsigned short)min(rx + rw, Wc); s_y1[i] = (unsigned short)min(ry + rh, Hc);
        s_g[i] = (unsigned char)(splitmix_draw(seed, b + 4) % 256ULL);
    }
    __syncthreads();
    const int p = blockIdx.x * 256 + threadIdx.x;
    if (p >= W * H) return;
    const int y = p / W, x = p - y * W;
    const int cx = 32 + dx + x, cy = 32 + dy + y;                 // canvas coordinates
    int v = -1;
    for (int i = R - 1; i >= 0; i--)                              // rectangles are painted in order: the last one covering the pixel wins
        if (cx >= s_x0[i] && cx < s_x1[i] && cy >= s_y0[i] && cy < s_y1[i]) { v = s_g[i]; break; }
    if (v < 0) {                                                  // smooth background: bilinear blend of the 16-px grid
        const int gy = cy >> 4, wy = cy & 15, gx = cx >> 4, wx = cx & 15;
        const int g00 = (int)(splitmix_draw(seed, (unsigned long long)gy * gw + gx + 1) % 256ULL);
        const int g01 = (int)(splitmix_draw(seed, (unsigned long long)gy * gw + gx + 2) % 256ULL);
        const int g10 = (int)(splitmix_draw(seed, (unsigned long long)(gy + 1) * gw + gx + 1) % 256ULL);
        const int g11 = (int)(splitmix_draw(seed, (unsigned long long)(gy + 1) * gw + gx + 2) % 256ULL);
        v = (g00 * (16 - wx) * (16 - wy) + g01 * wx * (16 - wy) + g10 * (16 - wx) * wy + g11 * wx * wy + 128) >> 8;
    }
    v += (int)(splitmix_draw(nseed, (unsigned long long)p + 1ULL) % 7ULL) - 3;
    out[(size_t)f * frame_pitch + p] = (uint8_t)min(max(v, 0), 255);
}

}  // namespace uvip

using namespace uvip;

extern "C" int uvip_synth_frames_device(const long long* d_seeds, const int* d_dxy, const long long* d_noise_seeds, int nframes, int w, int h,
                                        uint8_t* d_out, size_t frame_pitch, void* stream)
{
    UVIP_CHECK_ARG(d_seeds && d_dxy && d_noise_seeds && d_out && nframes >= 1 && w >= 16 && h >= 16 && frame_pitch >= (size_t)w * h);
    UVIP_CHECK_ARG(((long long)(w + 64) * (h + 64)) / 1500 <= SYN_MAXR && (long long)w * h < (1LL << 31));
    k_synth_frames<<<dim3((unsigned)div_up(w * h, 256), (unsigned)nframes), 256, 0, (cudaStream_t)stream>>>(d_seeds, d_dxy, d_noise_seeds, w, h, d_out, frame_pitch);
    UVIP_CUDA(cudaGetLastError());
    return UVIP_OK;
}
