// Which pipe does an integer op issue on?  Developer microbenchmark (sm_100a): throughput of op A alone, op B alone, and A+B
// interleaved; if time(A+B) ~= max(time(A), time(B)) the two ops use different pipes.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define OPS(asmtext)                                                               \
    _Pragma("unroll 1") for (int i = 0; i < ITERS; ++i) {                          \
        asm volatile(asmtext : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(b0), "+r"(b1), "+r"(b2), "+r"(b3) : "r"(k), "r"(m)); \
    }
template <int MODE>
__global__ void bench(unsigned *out, unsigned k, unsigned m, long long *cyc) {
    unsigned a0 = threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, b0 = a0 * 3, b1 = b0 + 1, b2 = b0 + 2, b3 = b0 + 3;
    long long t0 = clock64();
    if (MODE == 0) OPS("lop3.b32 %0,%0,%8,%9,0x96; lop3.b32 %1,%1,%8,%9,0x96; lop3.b32 %2,%2,%8,%9,0x96; lop3.b32 %3,%3,%8,%9,0x96;"
                       "lop3.b32 %4,%4,%8,%9,0x96; lop3.b32 %5,%5,%8,%9,0x96; lop3.b32 %6,%6,%8,%9,0x96; lop3.b32 %7,%7,%8,%9,0x96;")
    if (MODE == 1) OPS("dp4a.u32.u32 %0,%0,%8,%9; dp4a.u32.u32 %1,%1,%8,%9; dp4a.u32.u32 %2,%2,%8,%9; dp4a.u32.u32 %3,%3,%8,%9;"
                       "dp4a.u32.u32 %4,%4,%8,%9; dp4a.u32.u32 %5,%5,%8,%9; dp4a.u32.u32 %6,%6,%8,%9; dp4a.u32.u32 %7,%7,%8,%9;")
    if (MODE == 2) OPS("lop3.b32 %0,%0,%8,%9,0x96; dp4a.u32.u32 %4,%4,%8,%9; lop3.b32 %1,%1,%8,%9,0x96; dp4a.u32.u32 %5,%5,%8,%9;"
                       "lop3.b32 %2,%2,%8,%9,0x96; dp4a.u32.u32 %6,%6,%8,%9; lop3.b32 %3,%3,%8,%9,0x96; dp4a.u32.u32 %7,%7,%8,%9;")
    if (MODE == 3) OPS("mad.lo.u32 %0,%0,%8,%9; mad.lo.u32 %1,%1,%8,%9; mad.lo.u32 %2,%2,%8,%9; mad.lo.u32 %3,%3,%8,%9;"
                       "mad.lo.u32 %4,%4,%8,%9; mad.lo.u32 %5,%5,%8,%9; mad.lo.u32 %6,%6,%8,%9; mad.lo.u32 %7,%7,%8,%9;")
    if (MODE == 4) OPS("mad.lo.u32 %0,%0,%8,%9; dp4a.u32.u32 %4,%4,%8,%9; mad.lo.u32 %1,%1,%8,%9; dp4a.u32.u32 %5,%5,%8,%9;"
                       "mad.lo.u32 %2,%2,%8,%9; dp4a.u32.u32 %6,%6,%8,%9; mad.lo.u32 %3,%3,%8,%9; dp4a.u32.u32 %7,%7,%8,%9;")
    if (MODE == 5) OPS("lop3.b32 %0,%0,%8,%9,0x96; mad.lo.u32 %4,%4,%8,%9; lop3.b32 %1,%1,%8,%9,0x96; mad.lo.u32 %5,%5,%8,%9;"
                       "lop3.b32 %2,%2,%8,%9,0x96; mad.lo.u32 %6,%6,%8,%9; lop3.b32 %3,%3,%8,%9,0x96; mad.lo.u32 %7,%7,%8,%9;")
    if (MODE == 6) OPS("shf.l.wrap.b32 %0,%0,%8,%9; shf.l.wrap.b32 %1,%1,%8,%9; shf.l.wrap.b32 %2,%2,%8,%9; shf.l.wrap.b32 %3,%3,%8,%9;"
                       "shf.l.wrap.b32 %4,%4,%8,%9; shf.l.wrap.b32 %5,%5,%8,%9; shf.l.wrap.b32 %6,%6,%8,%9; shf.l.wrap.b32 %7,%7,%8,%9;")
    if (MODE == 7) OPS("prmt.b32 %0,%0,%8,%9; prmt.b32 %1,%1,%8,%9; prmt.b32 %2,%2,%8,%9; prmt.b32 %3,%3,%8,%9;"
                       "prmt.b32 %4,%4,%8,%9; prmt.b32 %5,%5,%8,%9; prmt.b32 %6,%6,%8,%9; prmt.b32 %7,%7,%8,%9;")
    if (MODE == 8) OPS("mad.hi.u32 %0,%0,%8,%9; mad.hi.u32 %1,%1,%8,%9; mad.hi.u32 %2,%2,%8,%9; mad.hi.u32 %3,%3,%8,%9;"
                       "mad.hi.u32 %4,%4,%8,%9; mad.hi.u32 %5,%5,%8,%9; mad.hi.u32 %6,%6,%8,%9; mad.hi.u32 %7,%7,%8,%9;")
    if (MODE == 9) OPS("add.u32 %0,%0,%8; add.u32 %1,%1,%8; add.u32 %2,%2,%8; add.u32 %3,%3,%8;"
                       "add.u32 %4,%4,%8; add.u32 %5,%5,%8; add.u32 %6,%6,%8; add.u32 %7,%7,%8;")
    if (MODE == 10) OPS("bfe.u32 %0,%0,%8,%9; bfe.u32 %1,%1,%8,%9; bfe.u32 %2,%2,%8,%9; bfe.u32 %3,%3,%8,%9;"
                        "bfe.u32 %4,%4,%8,%9; bfe.u32 %5,%5,%8,%9; bfe.u32 %6,%6,%8,%9; bfe.u32 %7,%7,%8,%9;")
    if (MODE == 11) OPS("shr.u32 %0,%0,%8; shr.u32 %1,%1,%8; shr.u32 %2,%2,%8; shr.u32 %3,%3,%8;"
                        "shr.u32 %4,%4,%8; shr.u32 %5,%5,%8; shr.u32 %6,%6,%8; shr.u32 %7,%7,%8;")
    if (MODE == 12) OPS("vadd.u32.u32.u32 %0,%0.b0,%8,%9; vadd.u32.u32.u32 %1,%1.b0,%8,%9; vadd.u32.u32.u32 %2,%2.b0,%8,%9; vadd.u32.u32.u32 %3,%3.b0,%8,%9;"
                        "vadd.u32.u32.u32 %4,%4.b0,%8,%9; vadd.u32.u32.u32 %5,%5.b0,%8,%9; vadd.u32.u32.u32 %6,%6.b0,%8,%9; vadd.u32.u32.u32 %7,%7.b0,%8,%9;")
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ b0 ^ b1 ^ b2 ^ b3;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(const char *name, unsigned *out, long long *cyc) {
    bench<MODE><<<148, 1024>>>(out, 3, 5, cyc);
    cudaDeviceSynchronize();
    long long h;
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // 8 warps per SMSP, 8 ops per iteration per warp
    printf("%-28s %8lld cycles  -> %.3f cycles per warp-instruction per SMSP\n", name, h, (double)h / (ITERS * 8.0 * 8.0));
}
int main() {
    unsigned *out; long long *cyc;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    run<0>("lop3", out, cyc); run<1>("dp4a", out, cyc); run<2>("lop3+dp4a", out, cyc); run<3>("imad", out, cyc);
    run<4>("imad+dp4a", out, cyc); run<5>("lop3+imad", out, cyc); run<6>("shf.wrap", out, cyc); run<7>("prmt", out, cyc);
    run<8>("imad.hi", out, cyc); run<9>("iadd", out, cyc); run<10>("bfe", out, cyc); run<11>("shr", out, cyc); run<12>("vadd.b0", out, cyc);
    return 0;
}
