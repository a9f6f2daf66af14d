// K1 (TMA variant) -- batched affine-coefficient assembly as a pipelined bulk-copy ring.
//
// Same reference chain as kernels_affine.cu (matvecmul! functions.jl:775-798 -> vecsubtract!/scale! :751-764,873-925 ->
// update!(::VectorAffineFunction) moi_interop.jl:64-81) for the common "simple" block: every touched output column is
// alpha x one source column.  The planner (pmb_plan.cu) merges adjacent columns into contiguous segments and packs them into
// chunks of <= 16 KB; persistent CTAs (3 per SM) stream the chunks through a 4-stage shared-memory ring with the bulk-copy engine
// (cp.async.bulk global->shared completing on an mbarrier, shared->global bulk_group; SASS UBLKCP), so HBM sees long
// back-to-back bursts with 48 KB in flight per CTA and no register staging.  alpha == 1 (the `C*x - d` block) is pure DMA
// issued by one warp; otherwise 128 threads scale the stage in place between the load and the store.
#include <cstdlib>

#include "pmb_kernels.h"

namespace pmb {

constexpr int kTmaMaxStages = 8;
constexpr int kTmaThreads = 128;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra LAB_DONE;\n"
        "bra LAB_WAIT;\n"
        "LAB_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ double alpha_at(const AlphaDev& a, int64_t inst) { return a.ptr ? a.c * a.ptr[inst * a.bstride] : a.c; }

// offsets d = const + sum alpha * v (same as kernels_affine.cu), run by the trailing CTAs of the launch
__device__ void tma_affine_offsets(const AffBlockDev& blk, int64_t batch, int ob, int nob) {
    if (!blk.d) return;
    const int64_t rb = (blk.d_rows + kTmaThreads - 1) / kTmaThreads;  // work item = (instance, block of kTmaThreads rows)
    for (int64_t w = ob; w < batch * rb; w += nob) {
        const int64_t inst = w / rb;
        const int64_t i = (w - inst * rb) * kTmaThreads + threadIdx.x;
        if (i >= blk.d_rows) continue;
        double acc = blk.d_const ? blk.d_const[i] : 0.0;
        for (int t = 0; t < blk.noffs; ++t) {
            const AffOffTermDev o = blk.offs[t];
            int64_t r = i - o.row_off;
            if (r >= 0 && r < o.rows) acc += alpha_at(o.alpha, inst) * o.src[inst * o.src_bstride + o.src_off + r];
        }
        blk.d[inst * blk.d_bstride + i] = acc;
    }
}

template <bool SCALE>
__global__ void __launch_bounds__(kTmaThreads) affine_assemble_tma(AffBlockDev blk, int64_t batch, int copy_blocks, int nstages) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full[kTmaMaxStages];
    // descriptors of the chunks in flight, one slot per stage (written by warp 0 before the load is issued)
    __shared__ int64_t d_inst[kTmaMaxStages];
    __shared__ int d_s0[kTmaMaxStages], d_s1[kTmaMaxStages];
    __shared__ uint32_t d_bytes[kTmaMaxStages];
    __shared__ AffSeg d_seg[kTmaMaxStages][32];
    if ((int)blockIdx.x >= copy_blocks) {
        tma_affine_offsets(blk, batch, blockIdx.x - copy_blocks, gridDim.x - copy_blocks);
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nstages; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const AffMatTermDev m = blk.mats[0];
    const int64_t total = (int64_t)blk.nchunks * batch;
    const int64_t first = blockIdx.x, stride = copy_blocks;
    const int64_t cnt = first < total ? (total - first + stride - 1) / stride : 0;

    // warp 0 issues every bulk copy; with SCALE the other warps only touch shared memory.
    // The static plan tables live in global memory; their (dependent) reads are kept off the issue path by fetching the
    // descriptor of a chunk one step before it is issued (and all prologue descriptors at once).
    struct Desc {
        int64_t inst;
        int s0, s1;
        uint32_t bytes;
        AffSeg seg;  // this lane's first segment of the chunk (the common chunk has <= 32 segments)
    };
    auto fetch = [&](int64_t i) {
        Desc d;
        const int64_t c = first + i * stride;
        d.inst = c / blk.nchunks;
        const int ch = (int)(c - d.inst * blk.nchunks);
        if (blk.uniform) {  // descriptors in arithmetic progression: no table reads on the issue path
            d.s0 = ch; d.s1 = ch + 1;
            d.bytes = (uint32_t)(ch + 1 == blk.nchunks ? blk.u_last_bytes : blk.u_bytes);
            d.seg = AffSeg{blk.u_src0 + ch * blk.u_src_stride, blk.u_dst0 + ch * blk.u_dst_stride, (int32_t)d.bytes, 0};
            return d;
        }
        d.s0 = blk.chunk_first[ch];
        d.s1 = blk.chunk_first[ch + 1];
        const AffSeg last = blk.segs[d.s1 - 1];
        d.bytes = (uint32_t)(last.stage_off + last.bytes);
        d.seg = blk.segs[d.s0 + lane < d.s1 ? d.s0 + lane : d.s1 - 1];
        return d;
    };
    auto issue_load = [&](int64_t i, const Desc& d) {
        const int st = (int)(i % nstages);
        unsigned char* stage = smem_raw + (size_t)st * kAffStageBytes;
        uint64_t* bar = &full[st];
        d_seg[st][lane] = d.seg;
        if (lane == 0) { d_inst[st] = d.inst; d_s0[st] = d.s0; d_s1[st] = d.s1; d_bytes[st] = d.bytes; }
        __syncwarp();
        if (lane == 0) mbar_expect_tx(bar, d.bytes);
        const double* src = m.src + d.inst * m.src_bstride;
        if (d.s0 + lane < d.s1) bulk_g2s(stage + d.seg.stage_off, src + d.seg.src_off, (uint32_t)d.seg.bytes, bar);
        for (int s = d.s0 + lane + 32; s < d.s1; s += 32) {
            const AffSeg sg = blk.segs[s];
            bulk_g2s(stage + sg.stage_off, src + sg.src_off, (uint32_t)sg.bytes, bar);
        }
    };
    auto issue_store = [&](int64_t i) {
        const int st = (int)(i % nstages);
        unsigned char* stage = smem_raw + (size_t)st * kAffStageBytes;
        const int s0 = d_s0[st], s1 = d_s1[st];
        double* dst = blk.C + d_inst[st] * blk.C_bstride;
        const AffSeg sg0 = d_seg[st][lane];
        if (s0 + lane < s1) bulk_s2g(dst + sg0.dst_off, stage + sg0.stage_off, (uint32_t)sg0.bytes);
        for (int s = s0 + lane + 32; s < s1; s += 32) {
            const AffSeg sg = blk.segs[s];
            bulk_s2g(dst + sg.dst_off, stage + sg.stage_off, (uint32_t)sg.bytes);
        }
        bulk_commit();
    };

    if (!SCALE && warp != 0) return;  // pure DMA: one warp drives the ring
    if (warp == 0) {
        // prologue: nstages - 1 loads in flight; descriptors are fetched four at a time so their table reads overlap
        for (int i0 = 0; i0 < nstages - 1 && i0 < cnt; i0 += 4) {
            Desc pro[4];
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < nstages - 1 && i0 + q < cnt) pro[q] = fetch(i0 + q);
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (i0 + q < nstages - 1 && i0 + q < cnt) issue_load(i0 + q, pro[q]);
        }
    }

    for (int64_t i = 0; i < cnt; ++i) {
        const uint32_t parity = (uint32_t)((i / nstages) & 1);
        if (SCALE) {
            mbar_wait(&full[i % nstages], parity);
            const int n2 = (int)(d_bytes[i % nstages] >> 4);
            const double a = alpha_at(m.alpha, d_inst[i % nstages]);
            double2* st = reinterpret_cast<double2*>(smem_raw + (size_t)(i % nstages) * kAffStageBytes);
            for (int j = threadIdx.x; j < n2; j += kTmaThreads) {
                double2 v = st[j];
                v.x *= a; v.y *= a;
                st[j] = v;
            }
            fence_async_smem();  // generic-proxy writes -> visible to the bulk-copy (async proxy) store
            __syncthreads();
        } else if (warp == 0) {
            mbar_wait(&full[i % nstages], parity);
        }
        if (warp == 0) {
            // descriptor of the chunk to issue at the end of this iteration: start its table reads now
            Desc nxt;
            const bool more = i + nstages - 1 < cnt;
            if (more) nxt = fetch(i + nstages - 1);
            issue_store(i);
            // refill the stage drained one iteration ago: its store must have finished READING shared memory
            bulk_wait_read<1>();
            __syncwarp();
            if (more) issue_load(i + nstages - 1, nxt);
        }
    }
    if (warp == 0) bulk_wait_all();
}

void configure_affine_tma_kernels() {
    cudaFuncSetAttribute(affine_assemble_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaMaxStages * kAffStageBytes);
    cudaFuncSetAttribute(affine_assemble_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTmaMaxStages * kAffStageBytes);
}

void launch_affine_assemble_tma(const AffBlockDev& blk, int64_t batch, int sm_count, cudaStream_t stream) {
    const int64_t total = (int64_t)blk.nchunks * batch;
    // prefer a grid that divides the chunk count so that every CTA moves the same number of chunks
    static int env_stages = -1, env_cps = -1;
    if (env_stages < 0) {
        const char* e1 = getenv("PMB_AFF_STAGES");
        const char* e2 = getenv("PMB_AFF_CTAS_PER_SM");
        env_stages = e1 ? atoi(e1) : 0;
        env_cps = e2 ? atoi(e2) : 0;
    }
    // measured at cfg2 (profiles/r01_c_affine_sweep.txt): the single issuing thread of a CTA is the limiter, so more CTAs per SM
    // with shallower rings beat one deep ring: 3 CTAs x 4 stages x 16 KB.
    // Long streams (cfg3: thousands of chunks per SM) are bandwidth- rather than issue-latency-bound: 2 CTAs x 6 stages measured
    // 6.0 TB/s there against 5.5 TB/s for 3 x 4.
    const bool long_stream = total > (int64_t)sm_count * 64;
    int cps = env_cps > 0 ? env_cps : (long_stream ? 2 : 3);
    int nstages = env_stages > 0 ? env_stages : (long_stream ? 6 : 4);
    if (nstages > kTmaMaxStages) nstages = kTmaMaxStages;
    if (nstages < 2) nstages = 2;
    int64_t cap = (int64_t)sm_count * cps;
    int64_t g = total < cap ? total : cap;
    if (total > cap)
        for (int64_t t = cap; t >= cap - cap / 8; --t)
            if (total % t == 0) { g = t; break; }
    const int64_t off_items = blk.d ? batch * ((blk.d_rows + kTmaThreads - 1) / kTmaThreads) : 0;
    const int off_blocks = (int)(off_items < 64 ? off_items : 64);
    dim3 grid((unsigned)(g + off_blocks));
    const size_t smem = (size_t)nstages * kAffStageBytes;
    if (blk.tma == 1) affine_assemble_tma<false><<<grid, kTmaThreads, smem, stream>>>(blk, batch, (int)g, nstages);
    else affine_assemble_tma<true><<<grid, kTmaThreads, smem, stream>>>(blk, batch, (int)g, nstages);
}

}  // namespace pmb
