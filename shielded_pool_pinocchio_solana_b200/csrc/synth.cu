// synth.cu -- on-device generation of synthetic curve points for the kernel sweep and the
// throughput benchmarks ("P_i = [k_i]G with k_i from SplitMix64 ...
// generate on device").  Not part of a proof; it exists so that 2^22..2^26-point base sets
// never have to be built by a CPU.
#include "capi.cuh"

namespace g16 {

__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9e3779b97f4a7c15ull;
    x = (x ^ (x >> 30)) * 0xbf58476d1ce4e5b9ull;
    x = (x ^ (x >> 27)) * 0x94d049bb133111ebull;
    return x ^ (x >> 31);
}

// k_i = 253-bit integer from four SplitMix64 outputs (always < r)
__host__ __device__ inline void synth_scalar(uint64_t seed, uint64_t i, uint32_t k[8]) {
    for (int w = 0; w < 4; w++) {
        uint64_t v = splitmix64(seed + 4 * i + w);
        k[2 * w] = (uint32_t)v;
        k[2 * w + 1] = (uint32_t)(v >> 32);
    }
    k[7] &= 0x1fffffffu;
}

template <class F>
__global__ void __launch_bounds__(128) k_synth_points(Affine<F> gen, uint64_t seed, uint32_t n, Affine<F>* out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t k[8];
    synth_scalar(seed, i, k);
    XYZZ<F> acc = XYZZ<F>::inf();
    for (int bit = 252; bit >= 0; bit--) {
        acc = acc.dbl();
        if ((k[bit >> 5] >> (bit & 31)) & 1u) acc.madd(gen);
    }
    Affine<F> a = acc.to_affine();
    // canonical (non-Montgomery) limbs on the way out
    Fp* c = reinterpret_cast<Fp*>(&a);
#pragma unroll
    for (int j = 0; j < (int)(sizeof(Affine<F>) / sizeof(Fp)); j++) c[j] = c[j].from_mont();
    out[i] = a;
}

static Fp fp_from_dec_limbs(const uint32_t l[8]) {
    Fp f;
    for (int i = 0; i < 8; i++) f.v[i] = l[i];
    return f.to_mont();
}

}  // namespace g16

using namespace g16;

extern "C" int g16_generate_points(g16_ctx* ctx, int g2, uint64_t seed, size_t n, uint8_t* out_be) {
    if (!ctx || !out_be || n == 0 || n > 0xffffffffull) {
        set_error("g16_generate_points: bad arguments");
        return G16_E_ARG;
    }
    G16_CUDA(cudaSetDevice(ctx->device));
    G16_LOCK(ctx);
    size_t ptsz = g2 ? sizeof(G2Affine) : sizeof(G1Affine);
    G16_TRY(ctx->results.ensure(ptsz * n));
    if (!g2) {
        G1Affine gen;
        gen.x = Fp::one();
        gen.y = Fp::one().dbl();
        k_synth_points<Fp><<<cdiv(n, 128), 128, 0, ctx->stream>>>(gen, seed, (uint32_t)n, (G1Affine*)ctx->results.ptr);
    } else {
        // gnark-crypto / EIP-197 G2 generator
        static const uint32_t X0[8] = {0xd992f6ed, 0x46debd5c, 0xf75edadd, 0x674322d4, 0x5e5c4479, 0x426a0066, 0x121f1e76, 0x1800deef};
        static const uint32_t X1[8] = {0xaef312c2, 0x97e485b7, 0x35a9e712, 0xf1aa4933, 0x31fb5d25, 0x7260bfb7, 0x920d483a, 0x198e9393};
        static const uint32_t Y0[8] = {0x66fa7daa, 0x4ce6cc01, 0x0c43d37b, 0xe3d1e769, 0x8dcb408f, 0x4aab7180, 0xdb8c6deb, 0x12c85ea5};
        static const uint32_t Y1[8] = {0xd122975b, 0x55acdadc, 0x70b38ef3, 0xbc4b3133, 0x690c3395, 0xec9e99ad, 0x585ff075, 0x090689d0};
        G2Affine gen;
        gen.x.c0 = fp_from_dec_limbs(X0); gen.x.c1 = fp_from_dec_limbs(X1);
        gen.y.c0 = fp_from_dec_limbs(Y0); gen.y.c1 = fp_from_dec_limbs(Y1);
        k_synth_points<Fp2><<<cdiv(n, 128), 128, 0, ctx->stream>>>(gen, seed, (uint32_t)n, (G2Affine*)ctx->results.ptr);
    }
    G16_CUDA(cudaGetLastError());
    std::vector<uint8_t> host(ptsz * n);
    G16_CUDA(cudaMemcpyAsync(host.data(), ctx->results.ptr, ptsz * n, cudaMemcpyDeviceToHost, ctx->stream));
    G16_CUDA(cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < n; i++) {
        if (!g2) g1_to_be(*reinterpret_cast<G1Affine*>(host.data() + ptsz * i), out_be + 64 * i);
        else g2_to_be(*reinterpret_cast<G2Affine*>(host.data() + ptsz * i), out_be + 128 * i);
    }
    return G16_OK;
}
