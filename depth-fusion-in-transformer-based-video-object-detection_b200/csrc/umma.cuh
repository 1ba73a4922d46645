// Thin inline-PTX wrappers for the Blackwell (sm_100a) tensor-core path: tcgen05.mma with
// accumulators in tensor memory (TMEM), shared-memory matrix descriptors, mbarriers, cp.async.
// Only what csrc/ffn_fused.cu needs; bit layouts follow the PTX ISA "tcgen05" matrix / instruction
// descriptor tables.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace umma {

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

// ---- K-major SWIZZLE_128B operand tile: [rows][64 x 16-bit] per k-block, 8-row groups of 1024 B,
// the 16-byte chunk index XOR-ed with (row % 8).  Byte offset of 16-byte chunk `c` (0..4*8-1 over a
// 256-wide K: k-block = c / 8) of row `r` inside an operand with `rows` rows per k-block.
__device__ __forceinline__ unsigned sw128_offset(int r, int c, int rows)
{
    const int kb = c >> 3, cc = c & 7;
    return (unsigned)(kb * rows * 128 + (r >> 3) * 1024 + (r & 7) * 128 + ((cc ^ (r & 7)) << 4));
}

// shared-memory matrix descriptor: K-major, SWIZZLE_128B, stride between 8-row groups 1024 B
__device__ __forceinline__ unsigned long long make_desc_sw128(const void* smem_ptr)
{
    const unsigned addr = smem_u32(smem_ptr);
    unsigned long long d = 0;
    d |= (unsigned long long)((addr & 0x3FFFF) >> 4);            // start address      bits [0,14)
    d |= (unsigned long long)1 << 16;                            // leading byte offset (unused for swizzled K-major)
    d |= (unsigned long long)(1024 >> 4) << 32;                  // stride byte offset  bits [32,46)
    d |= (unsigned long long)1 << 46;                            // descriptor version (sm_100)
    d |= (unsigned long long)2 << 61;                            // layout: SWIZZLE_128B
    return d;
}

// the same descriptor `bytes` further into the tile (only the 14-bit start-address field changes)
__device__ __forceinline__ unsigned long long desc_advance(unsigned long long desc, unsigned bytes)
{
    return desc + (unsigned long long)(bytes >> 4);
}

// one lane of a fully converged warp
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// instruction descriptor, kind::f16: D fp32, A/B bf16, both K-major, M x N tile
__device__ __forceinline__ unsigned make_idesc_bf16(int m, int n)
{
    return (1u << 4) | (1u << 7) | (1u << 10) | ((unsigned)(n >> 3) << 17) | ((unsigned)(m >> 4) << 24);
}

__device__ __forceinline__ void mma_bf16(unsigned tmem_d, unsigned long long desc_a, unsigned long long desc_b,
                                         unsigned idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"((unsigned)accumulate) : "memory");
}

// A operand from tensor memory (M lanes x K 16-bit elements packed two per 32-bit column), B from shared memory
__device__ __forceinline__ void mma_bf16_ts(unsigned tmem_d, unsigned tmem_a, unsigned long long desc_b,
                                            unsigned idesc, bool accumulate)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}\n"
        :: "r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"((unsigned)accumulate) : "memory");
}

// shared memory -> tensor memory: 128 rows x 256 bits (16 bf16 = 8 columns per lane), source described like
// the A operand of one K=16 MMA step
__device__ __forceinline__ void tmem_cp_128x256b(unsigned tmem_dst, unsigned long long desc_src)
{
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" :: "r"(tmem_dst), "l"(desc_src) : "memory");
}

// all previously issued tcgen05.mma of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void mma_commit(unsigned long long* bar)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 :: "r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// one warp allocates `cols` (power of two >= 32) TMEM columns; the base address lands in *dst (shared)
__device__ __forceinline__ void tmem_alloc(unsigned* dst, int cols)
{
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" :: "r"(smem_u32(dst)), "r"(cols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(unsigned base, int cols)
{
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" :: "r"(base), "r"(cols) : "memory");
}

// 32 consecutive fp32 columns of this thread's TMEM lane (lane = 32 * (warp % 4) + laneid)
__device__ __forceinline__ void tmem_ld32(unsigned taddr, float* v)
{
    unsigned r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 consecutive fp32 columns of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld16(unsigned taddr, float* v)
{
    unsigned r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}\n"
        :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// named barrier among `threads` threads (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int threads)
{
    asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(threads) : "memory");
}

// ---- TMA: one 2-D box global -> shared (SWIZZLE_128B box = the UMMA K-major SWIZZLE_128B tile) ----
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const void* tensor_map, int c0, int c1,
                                            unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        :: "r"(smem_u32(smem_dst)), "l"(tensor_map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
// the same box, only as far as L2 (no shared-memory destination, no barrier): hides HBM latency of a later load
__device__ __forceinline__ void tma_prefetch_l2_2d(const void* tensor_map, int c0, int c1)
{
    asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global [%0, {%1, %2}];" :: "l"(tensor_map), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const void* tensor_map)
{
    asm volatile("prefetch.tensormap [%0];" :: "l"(tensor_map) : "memory");
}

// ---- thread-block clusters: rank, cluster barrier, TMA multicast, tcgen05.commit onto every CTA's barrier ----
__device__ __forceinline__ unsigned cluster_ctarank()
{
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_multicast(void* smem_dst, const void* tensor_map, int c0, int c1,
                                                      unsigned long long* bar, unsigned short cta_mask)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        :: "r"(smem_u32(smem_dst)), "l"(tensor_map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(cta_mask) : "memory");
}
__device__ __forceinline__ void mma_commit_multicast(unsigned long long* bar, unsigned short cta_mask)
{
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 :: "r"(smem_u32(bar)), "h"(cta_mask) : "memory");
}

// one 2-D box shared -> global (bulk group); rows / columns past the end of the tensor are clipped
__device__ __forceinline__ void tma_store_2d(const void* tensor_map, const void* smem_src, int c0, int c1)
{
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 :: "l"(tensor_map), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// the bulk groups committed so far have READ their shared-memory source (it may be rewritten)
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ---- generic-proxy writes -> async-proxy (tensor core) reads ----
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

}  // namespace umma
