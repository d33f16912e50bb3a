// k_migrate.cu -- multi-GPU half of stage 5: particles whose ancestor lives on
// another rank migrate with their page table and sub-tiles.
//
// Every rank runs the identical global resample (k_resample.cu), so both ends
// of every transfer know which particles move; no negotiation is needed.  For
// one destination peer the sender
//   migrate_claim   de-duplicates the sub-tiles referenced by the departing
//                   particles (a sub-tile shared by several of them crosses
//                   NVLink once) and assigns compact indices,
//   migrate_pack    writes [records | page tables with compact indices | tile
//                   payloads] into one contiguous device buffer (moved by NCCL),
// and the receiver
//   migrate_unpack  adopts the payloads into fresh pool sub-tiles and points the
//                   destination slots' page tables at them (refcount = number of
//                   local descendants).
// Replaces Robot.copy / HybridMap.copy (robot.py:141-149, hybridmap.py:315-320)
// for descendants on another GPU.
#include "common.cuh"

#define MG_REC_DOUBLES 13            // pose 3 + cov 9 + exists mask (bit pattern)

__host__ __device__ inline size_t mg_rec_bytes(int nsub) { return ((size_t)MG_REC_DOUBLES * 8 + (size_t)nsub * 4 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t mg_header_bytes(int n, int nsub) { return (mg_rec_bytes(nsub) * (size_t)n + 255) & ~(size_t)255; }

// pass 1: every allocated page-table entry of the departing particles claims its sub-tile once
__global__ void migrate_claim_kernel(RbCtx c, const int *__restrict__ slots, int n, uint32_t *mark, uint32_t *list,
                                     int *count)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const uint32_t *pt = c.pt + (size_t)slots[warp] * c.nsub;
    for (int e = lane; e < c.nsub; e += 32) {
        uint32_t t = pt[e];
        if (t == RB_NONE) continue;
        if (atomicCAS(&mark[t], RB_NONE, 0xFFFFFFFEu) == RB_NONE) {
            int idx = atomicAdd(count, 1);
            list[idx] = t;
        }
    }
}

// pass 2a: compact index of every claimed sub-tile
__global__ void migrate_index_kernel(uint32_t *mark, const uint32_t *__restrict__ list, const int *count)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < *count) mark[list[i]] = (uint32_t)i;
}

// pass 2b: records + translated page tables
__global__ void migrate_pack_records_kernel(RbCtx c, const int *__restrict__ slots, int n, const uint32_t *__restrict__ mark,
                                            unsigned char *buf)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= n) return;
    const int s = slots[warp];
    unsigned char *rec = buf + mg_rec_bytes(c.nsub) * (size_t)warp;
    double *d = reinterpret_cast<double *>(rec);
    if (lane < 3) d[lane] = c.pose[3 * (size_t)s + lane];
    if (lane < 9) d[3 + lane] = c.cov[9 * (size_t)s + lane];
    if (lane == 0) d[12] = __longlong_as_double((long long)c.exists[s]);
    uint32_t *pt_out = reinterpret_cast<uint32_t *>(rec + MG_REC_DOUBLES * 8);
    const uint32_t *pt = c.pt + (size_t)s * c.nsub;
    for (int e = lane; e < c.nsub; e += 32) {
        uint32_t t = pt[e];
        pt_out[e] = t == RB_NONE ? RB_NONE : mark[t];
    }
}

// pass 2c: tile payloads (one CTA per sub-tile), then release the claim
__global__ void __launch_bounds__(256) migrate_pack_tiles_kernel(RbCtx c, uint32_t *mark, const uint32_t *__restrict__ list,
                                                                 const int *count, unsigned char *payload)
{
    const int i = blockIdx.x;
    if (i >= *count) return;
    const uint32_t t = list[i];
    const uint4 *src = reinterpret_cast<const uint4 *>(c.pool + (size_t)t * RB_SUB_BYTES);
    uint4 *dst = reinterpret_cast<uint4 *>(payload + (size_t)i * RB_SUB_BYTES);
    for (int q = threadIdx.x; q < RB_SUB_BYTES / 16; q += blockDim.x) dst[q] = src[q];
    if (threadIdx.x == 0) mark[t] = RB_NONE;
}

// receiver: fresh sub-tiles for the payloads; map[i] = new pool index
__global__ void __launch_bounds__(256) migrate_adopt_tiles_kernel(RbCtx c, const unsigned char *__restrict__ payload,
                                                                  int n_tiles, uint32_t *map)
{
    __shared__ uint32_t s_t;
    const int i = blockIdx.x;
    if (i >= n_tiles) return;
    if (threadIdx.x == 0) {
        int idx = atomicSub(c.free_count, 1) - 1;
        if (idx < 0) { atomicAdd(c.free_count, 1); atomicExch(&c.flags->pool_exhausted, 1); s_t = RB_NONE; }
        else { s_t = c.free_list[idx]; c.refcnt[s_t] = 0u; }
        map[i] = s_t;
    }
    __syncthreads();
    const uint32_t t = s_t;
    if (t == RB_NONE) return;
    const uint4 *src = reinterpret_cast<const uint4 *>(payload + (size_t)i * RB_SUB_BYTES);
    uint4 *dst = reinterpret_cast<uint4 *>(c.pool + (size_t)t * RB_SUB_BYTES);
    for (int q = threadIdx.x; q < RB_SUB_BYTES / 16; q += blockDim.x) dst[q] = src[q];
}

// receiver: destination slot j takes received record rec_idx[j]
__global__ void migrate_place_kernel(RbCtx c, const unsigned char *__restrict__ buf, const int *__restrict__ dst_slots,
                                     const int *__restrict__ rec_idx, int m, const uint32_t *__restrict__ map)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= m) return;
    const int j = dst_slots[warp];
    const unsigned char *rec = buf + mg_rec_bytes(c.nsub) * (size_t)rec_idx[warp];
    const double *d = reinterpret_cast<const double *>(rec);
    if (lane < 3) c.pose2[3 * (size_t)j + lane] = d[lane];
    if (lane < 9) c.cov2[9 * (size_t)j + lane] = d[3 + lane];
    if (lane == 0) {
        c.exists2[j] = (unsigned long long)__double_as_longlong(d[12]);
        c.weight[j] = 1.0;                                                   // main.py:77-78
    }
    const uint32_t *pt_in = reinterpret_cast<const uint32_t *>(rec + MG_REC_DOUBLES * 8);
    uint32_t *dst = c.pt2 + (size_t)j * c.nsub;
    for (int e = lane; e < c.nsub; e += 32) {
        uint32_t ci = pt_in[e];
        uint32_t t = ci == RB_NONE ? RB_NONE : map[ci];
        dst[e] = t;
        if (t != RB_NONE) atomicAdd(&c.refcnt[t], 1u);
    }
}

void rb_launch_migrate_claim(const RbCtx &c, const int *slots_dev, int n, uint32_t *mark, uint32_t *list, int *count,
                             cudaStream_t s)
{
    cudaMemsetAsync(count, 0, sizeof(int), s);
    if (n > 0) migrate_claim_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(c, slots_dev, n, mark, list, count);
}

void rb_launch_migrate_pack(const RbCtx &c, const int *slots_dev, int n, int n_tiles, uint32_t *mark, uint32_t *list,
                            int *count, unsigned char *buf, cudaStream_t s)
{
    if (n <= 0) return;
    if (n_tiles > 0) migrate_index_kernel<<<(n_tiles + 255) / 256, 256, 0, s>>>(mark, list, count);
    migrate_pack_records_kernel<<<(n * 32 + 255) / 256, 256, 0, s>>>(c, slots_dev, n, mark, buf);
    if (n_tiles > 0)
        migrate_pack_tiles_kernel<<<n_tiles, 256, 0, s>>>(c, mark, list, count, buf + mg_header_bytes(n, c.nsub));
}

void rb_launch_migrate_unpack(const RbCtx &c, const unsigned char *buf, int n, int n_tiles, const int *dst_slots_dev,
                              const int *rec_idx_dev, int m, uint32_t *map, cudaStream_t s)
{
    if (n_tiles > 0) migrate_adopt_tiles_kernel<<<n_tiles, 256, 0, s>>>(c, buf + mg_header_bytes(n, c.nsub), n_tiles, map);
    if (m > 0) migrate_place_kernel<<<(m * 32 + 255) / 256, 256, 0, s>>>(c, buf, dst_slots_dev, rec_idx_dev, m, map);
}

// ---- pull over peer memory -----------------------------------------------------------
// The receiver reads the source ranks' page tables, particle state and sub-tiles
// directly through NVLink-mapped pointers (CUDA IPC) and writes them straight into
// their final place: no pack, no staging buffer, no size negotiation, and -- because
// the plan is the ancestor vector already on the device -- no host synchronisation.
// The sources' buffers are read-only between their last stage-4 kernel and the
// job-wide barrier that follows the pulls (thesis_b200/dist.py).
// mark[r * pool_tiles + t] de-duplicates sub-tile t of rank r; list/list_rank hold the claims.

__device__ __forceinline__ bool pull_source(const RbCtx &c, int j, int &r, int &s)
{
    if (c.flags->resample_error || !c.flags->did_resample) return false;     // particles unchanged
    const int a = c.ancestors[c.rank * c.N + j];
    r = a / c.N;
    s = a - r * c.N;
    return r != c.rank;
}

// pass 1: every allocated entry of a needed remote particle claims its remote sub-tile once
__global__ void pull_claim_kernel(RbCtx c, RbPeers peers, uint32_t *mark, uint32_t *list, unsigned char *list_rank, int *count)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= c.N) return;
    int r, s;
    if (!pull_source(c, warp, r, s)) return;
    const uint32_t *pt = peers.p[r].pt + (size_t)s * c.nsub;
    uint32_t *mk = mark + (size_t)r * c.pool_tiles;
    uint32_t *raw_local = c.pt2 + (size_t)warp * c.nsub;                     // the slot's page table: remote indices for now
    // eight remote loads in flight per lane (a load over NVLink takes microseconds), then the claims
    for (int e0 = 0; e0 < c.nsub; e0 += 32 * 8) {
        uint32_t t[8];
#pragma unroll
        for (int k = 0; k < 8; k++) { const int e = e0 + 32 * k + lane; t[k] = e < c.nsub ? pt[e] : RB_NONE; }
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const int e = e0 + 32 * k + lane;
            if (e >= c.nsub) continue;
            raw_local[e] = t[k];
            if (t[k] == RB_NONE || t[k] >= c.pool_tiles) continue;
            if (atomicCAS(&mk[t[k]], RB_NONE, 0xFFFFFFFEu) == RB_NONE) {
                const int idx = atomicAdd(count, 1);
                if ((uint32_t)idx < c.pool_tiles) { list[idx] = t[k]; list_rank[idx] = (unsigned char)r; }
                else atomicExch(&c.flags->pool_exhausted, 1);
            }
        }
    }
}

// pass 2a: a fresh local sub-tile for every claimed remote one; mark = local index, list_local = the same by claim
__global__ void __launch_bounds__(256) pull_alloc_kernel(RbCtx c, uint32_t *mark, const uint32_t *__restrict__ list,
                                                         const unsigned char *__restrict__ list_rank, uint32_t *list_local, const int *count)
{
    const int n = min(*count, (int)c.pool_tiles);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t t;
        const int idx = atomicSub(c.free_count, 1) - 1;
        if (idx < 0) { atomicAdd(c.free_count, 1); atomicExch(&c.flags->pool_exhausted, 1); t = RB_NONE; }
        else { t = c.free_list[idx]; c.refcnt[t] = 0u; }
        list_local[i] = t;
        mark[(size_t)list_rank[i] * c.pool_tiles + list[i]] = t == RB_NONE ? 0xFFFFFFFDu : t;
    }
}

// pass 2b: the payloads, copied across the link (grid-stride over the device-side count).  Nothing on this rank reads the
// new sub-tiles before the particles that own them are matched, so this kernel may run on a side stream beside the
// rest of the resample and the next scan's matching of the local particles (rbpf_migrate_pull_async).  Measured at 4 GPUs:
// no gain (6.51 against 6.52 ms per scan) -- beside two resident CTAs of match_kernel an SM has 4,096 registers and no
// shared memory left, so the copy takes CTA slots from the matcher for as long as it runs, and the migrated particles are
// matched in a short second launch with a poor tail; a copy kernel small enough to fit beside the matcher (128 threads x
// 32 registers per SM) was slower still (6.60 ms).  thesis_b200.dist therefore pulls in stream order unless
// RBPF_DIST_OVERLAP=1.
__global__ void __launch_bounds__(256) pull_copy_kernel(RbCtx c, RbPeers peers, const uint32_t *__restrict__ list,
                                                        const unsigned char *__restrict__ list_rank,
                                                        const uint32_t *__restrict__ list_local, const int *count)
{
    const int n = min(*count, (int)c.pool_tiles);
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        const uint32_t rt = list[i], t = list_local[i];
        const int r = list_rank[i];
        if (t == RB_NONE) continue;
        const uint4 *src = reinterpret_cast<const uint4 *>(peers.p[r].pool + (size_t)rt * RB_SUB_BYTES);
        uint4 *dst = reinterpret_cast<uint4 *>(c.pool + (size_t)t * RB_SUB_BYTES);
        uint4 v[7];                                                          // 1,600 x 16 B over 256 threads: all loads first
#pragma unroll
        for (int k = 0; k < 7; k++) { const int q = threadIdx.x + 256 * k; if (q < RB_SUB_BYTES / 16) v[k] = src[q]; }
#pragma unroll
        for (int k = 0; k < 7; k++) { const int q = threadIdx.x + 256 * k; if (q < RB_SUB_BYTES / 16) dst[q] = v[k]; }
    }
}

// pass 3: a local slot with a remote ancestor becomes a copy of that particle
__global__ void pull_place_kernel(RbCtx c, RbPeers peers, const uint32_t *__restrict__ mark)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= c.N) return;
    int r, s;
    if (!pull_source(c, warp, r, s)) return;
    const int j = warp;
    const RbPeer &peer = peers.p[r];
    if (lane < 3) c.pose2[3 * (size_t)j + lane] = peer.pose[3 * (size_t)s + lane];
    if (lane < 9) c.cov2[9 * (size_t)j + lane] = peer.cov[9 * (size_t)s + lane];
    if (lane == 0) {
        c.exists2[j] = peer.exists[s];
        c.weight[j] = 1.0;                                                   // main.py:77-78
        c.pulled[j] = 1;                                                     // its sub-tiles may still be on their way
    }
    const uint32_t *mk = mark + (size_t)r * c.pool_tiles;
    uint32_t *dst = c.pt2 + (size_t)j * c.nsub;                              // holds the remote indices (pull_claim_kernel)
    for (int e = lane; e < c.nsub; e += 32) {
        const uint32_t rt = dst[e];
        uint32_t t = RB_NONE;
        if (rt != RB_NONE && rt < c.pool_tiles) {
            t = mk[rt];
            if (t >= c.pool_tiles) t = RB_NONE;                              // pool exhausted (flag is set)
        }
        dst[e] = t;
        if (t != RB_NONE) atomicAdd(&c.refcnt[t], 1u);
    }
}

// pass 4: release the claims
__global__ void pull_release_kernel(RbCtx c, uint32_t *mark, const uint32_t *__restrict__ list,
                                    const unsigned char *__restrict__ list_rank, const int *count)
{
    const int n = min(*count, (int)c.pool_tiles);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
        mark[(size_t)list_rank[i] * c.pool_tiles + list[i]] = RB_NONE;
}

// copy_stream: where the payload copies run (the caller has made it wait for `s` up to the allocation through ev_alloc);
// the same stream as `s` keeps everything in order.
void rb_launch_migrate_pull(const RbCtx &c, const RbPeers &peers, uint32_t *mark, uint32_t *list, unsigned char *list_rank,
                            uint32_t *list_local, int *count, cudaStream_t s, cudaStream_t copy_stream, cudaEvent_t ev_alloc)
{
    const int wblocks = (c.N * 32 + 255) / 256;
    cudaMemsetAsync(count, 0, sizeof(int), s);
    cudaMemsetAsync(c.pulled, 0, (size_t)c.N, s);
    pull_claim_kernel<<<wblocks, 256, 0, s>>>(c, peers, mark, list, list_rank, count);
    pull_alloc_kernel<<<148, 256, 0, s>>>(c, mark, list, list_rank, list_local, count);
    if (copy_stream != s) {
        cudaEventRecord(ev_alloc, s);
        cudaStreamWaitEvent(copy_stream, ev_alloc, 0);
    }
    pull_copy_kernel<<<148 * 8, 256, 0, copy_stream>>>(c, peers, list, list_rank, list_local, count);
    pull_place_kernel<<<wblocks, 256, 0, s>>>(c, peers, mark);
    pull_release_kernel<<<148, 256, 0, s>>>(c, mark, list, list_rank, count);
}

size_t rb_migrate_bytes(int n, int n_tiles, int nsub) { return mg_header_bytes(n, nsub) + (size_t)n_tiles * RB_SUB_BYTES; }
