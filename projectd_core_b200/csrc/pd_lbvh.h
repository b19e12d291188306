/*
 * pd_lbvh.h -- the ray caster's bounding-volume hierarchy BUILT ON THE DEVICE ("BVH build on device").
 *
 * The reference has no tree of its own (its rays go through ODE / OPCODE, RayCasterODE.cpp); the product's general rays (pd_track.h
 * ray_cast: Track::computeFatPoints' traces, pd_raycast) walk a binary tree over the track's triangles.  load_track builds one on the
 * host by recursive median splits; for very large tracks (BASELINE configs[3]: 1 M triangles) the same structure is built here in a
 * few launches as a linear BVH (Lauritzen / Karras 2012):
 *
 *   1. k_lbvh_keys     centroid of every triangle -> 30-bit Morton code in the track's bounding box; key = code << 32 | triangle index
 *                      (unique keys: no special case for equal codes)
 *   2. cub radix sort  of the 64-bit keys
 *   3. k_lbvh_internal one thread per internal node: its key range and split from the longest common prefixes of neighbouring keys
 *   4. k_lbvh_refit    one thread per leaf climbs towards the root; the second arrival at a node (atomic counter) unites the boxes
 *   5. k_lbvh_emit     the product's node format (pd_track.h BvhNode: children adjacent, `left` / `left + 1`): the two children of internal
 *                      node i are written to slots 1 + 2 i and 2 + 2 i, the root to slot 0; a leaf holds ONE triangle (its index in the
 *                      order of `tris` / `triRaw`, which is not changed)
 *   6. k_lbvh_depth    the deepest leaf: the traversal stack of ray_cast holds depth + 1 entries, so a tree deeper than 46 is refused
 *                      (the caller keeps the host-built tree)
 *
 * Closest-hit rays do not depend on the shape of the tree: tests/test_gpu_parity.py::test_device_bvh_rays_equal_host_bvh_rays.
 */
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include "pd_track.h"

namespace pd {

__device__ __forceinline__ uint32_t lbvh_expand10(uint32_t v) {       /* 10 bits -> every third bit */
    v = (v * 0x00010001u) & 0xFF0000FFu; v = (v * 0x00000101u) & 0x0F00F00Fu; v = (v * 0x00000011u) & 0xC30C30C3u; v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__global__ void k_lbvh_keys(const float* __restrict__ triRaw, int n, float3 lo, float3 inv, unsigned long long* __restrict__ keys) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* p = triRaw + (size_t)i * 9;
    const float cx = (p[0] + p[3] + p[6]) * (1.0f / 3.0f), cy = (p[1] + p[4] + p[7]) * (1.0f / 3.0f), cz = (p[2] + p[5] + p[8]) * (1.0f / 3.0f);
    const uint32_t x = (uint32_t)fminf(fmaxf((cx - lo.x) * inv.x * 1024.0f, 0.0f), 1023.0f);
    const uint32_t y = (uint32_t)fminf(fmaxf((cy - lo.y) * inv.y * 1024.0f, 0.0f), 1023.0f);
    const uint32_t z = (uint32_t)fminf(fmaxf((cz - lo.z) * inv.z * 1024.0f, 0.0f), 1023.0f);
    const uint32_t code = (lbvh_expand10(x) << 2) | (lbvh_expand10(z) << 1) | lbvh_expand10(y);
    keys[i] = ((unsigned long long)code << 32) | (unsigned long long)(uint32_t)i;
}
__device__ __forceinline__ int lbvh_delta(const unsigned long long* __restrict__ keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll((long long)(keys[i] ^ keys[j]));
}
/* children[2 i], children[2 i + 1]: >= 0 internal node index, < 0: ~leaf (position in the sorted key array) */
__global__ void k_lbvh_internal(const unsigned long long* __restrict__ keys, int n, int* __restrict__ children, int* __restrict__ parentInternal, int* __restrict__ parentLeaf) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1) if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1; ; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t == 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int first = min(i, j), last = max(i, j);
    const int left = (first == gamma) ? ~gamma : gamma, right = (last == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    children[2 * i] = left; children[2 * i + 1] = right;
    if (left >= 0) parentInternal[left] = i; else parentLeaf[~left] = i;
    if (right >= 0) parentInternal[right] = i; else parentLeaf[~right] = i;
    if (i == 0) parentInternal[0] = -1;
}
__device__ __forceinline__ void lbvh_tri_box(const float* __restrict__ triRaw, int t, float* mn, float* mx) {
    const float* p = triRaw + (size_t)t * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) { mn[k] = fminf(p[k], fminf(p[3 + k], p[6 + k])); mx[k] = fmaxf(p[k], fmaxf(p[3 + k], p[6 + k])); }
}
__global__ void k_lbvh_refit(const unsigned long long* __restrict__ keys, const float* __restrict__ triRaw, int n, const int* __restrict__ children, const int* __restrict__ parentInternal,
                             const int* __restrict__ parentLeaf, int* __restrict__ visits, float* __restrict__ boxes /* [n - 1][6] */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parentLeaf[i];
    while (node >= 0) {
        if (atomicAdd(&visits[node], 1) == 0) return;          /* the first child to arrive leaves; the second finds both boxes written */
        __threadfence();
        float mn[3] = {3.4e38f, 3.4e38f, 3.4e38f}, mx[3] = {-3.4e38f, -3.4e38f, -3.4e38f};
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int ch = children[2 * node + c];
            float a[3], b[3];
            if (ch < 0) lbvh_tri_box(triRaw, (int)(uint32_t)(keys[~ch] & 0xffffffffull), a, b);
            else { const volatile float* q = boxes + (size_t)ch * 6; a[0] = q[0]; a[1] = q[1]; a[2] = q[2]; b[0] = q[3]; b[1] = q[4]; b[2] = q[5]; }
#pragma unroll
            for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], a[k]); mx[k] = fmaxf(mx[k], b[k]); }
        }
        float* q = boxes + (size_t)node * 6; q[0] = mn[0]; q[1] = mn[1]; q[2] = mn[2]; q[3] = mx[0]; q[4] = mx[1]; q[5] = mx[2];
        __threadfence();
        node = parentInternal[node];
    }
}
__device__ __forceinline__ void lbvh_write_node(BvhNode* out, const float* mn, const float* mx, int left, int count) {
    BvhNode nd;
#pragma unroll
    for (int k = 0; k < 3; ++k) {      /* the host builder's padding: the box test is only a filter */
        const float pad = 1e-4f * fmaxf(1.0f, fmaxf(fabsf(mn[k]), fabsf(mx[k])));
        nd.bmin[k] = mn[k] - pad; nd.bmax[k] = mx[k] + pad;
    }
    nd.left = left; nd.count = count;
    *out = nd;
}
__global__ void k_lbvh_emit(const unsigned long long* __restrict__ keys, const float* __restrict__ triRaw, int n, const int* __restrict__ children, const float* __restrict__ boxes, BvhNode* __restrict__ nodes) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    if (i == 0) lbvh_write_node(nodes, boxes, boxes + 3, 1, 0);
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int ch = children[2 * i + c];
        BvhNode* slot = nodes + 1 + 2 * (size_t)i + c;
        if (ch < 0) { const int t = (int)(uint32_t)(keys[~ch] & 0xffffffffull); float a[3], b[3]; lbvh_tri_box(triRaw, t, a, b); lbvh_write_node(slot, a, b, t, 1); }
        else lbvh_write_node(slot, boxes + (size_t)ch * 6, boxes + (size_t)ch * 6 + 3, 1 + 2 * ch, 0);
    }
}
__global__ void k_lbvh_depth(int n, const int* __restrict__ parentInternal, const int* __restrict__ parentLeaf, int* __restrict__ maxDepth) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int d = 0;
    for (int node = parentLeaf[i]; node >= 0; node = parentInternal[node]) ++d;
    atomicMax(maxDepth, d);
}

/* Builds the tree over the n triangles of dTriRaw (device, 9 floats each) inside the box [lo, hi]; on success *outNodes (device, 2 n - 1 nodes,
 * cudaMalloc'ed) replaces the host-built tree.  Returns 0, or a message. */
inline const char* lbvh_build_on_device(const float* dTriRaw, int n, const float lo[3], const float hi[3], cudaStream_t stream, BvhNode** outNodes, int* outNodeCount, int* outDepth) {
    if (n < 2) return "fewer than two triangles";
    unsigned long long *keysA = nullptr, *keysB = nullptr; int *children = nullptr, *parentI = nullptr, *parentL = nullptr, *visits = nullptr, *dDepth = nullptr; float* boxes = nullptr; void* tmp = nullptr; BvhNode* nodes = nullptr;
    size_t tmpBytes = 0;
    const char* err = nullptr;
    auto ok = [&](cudaError_t e) { if (e != cudaSuccess && !err) err = cudaGetErrorString(e); return e == cudaSuccess; };
    do {
        if (!ok(cudaMalloc(&keysA, (size_t)n * 8)) || !ok(cudaMalloc(&keysB, (size_t)n * 8)) || !ok(cudaMalloc(&children, (size_t)(n - 1) * 8)) || !ok(cudaMalloc(&parentI, (size_t)(n - 1) * 4)) ||
            !ok(cudaMalloc(&parentL, (size_t)n * 4)) || !ok(cudaMalloc(&visits, (size_t)(n - 1) * 4)) || !ok(cudaMalloc(&boxes, (size_t)(n - 1) * 24)) || !ok(cudaMalloc(&dDepth, 4)) ||
            !ok(cudaMalloc(&nodes, (size_t)(2 * n - 1) * sizeof(BvhNode)))) break;
        const float3 l3 = make_float3(lo[0], lo[1], lo[2]);
        const float3 inv = make_float3(hi[0] > lo[0] ? 1.0f / (hi[0] - lo[0]) : 0.0f, hi[1] > lo[1] ? 1.0f / (hi[1] - lo[1]) : 0.0f, hi[2] > lo[2] ? 1.0f / (hi[2] - lo[2]) : 0.0f);
        const int B = 256, G = (n + B - 1) / B;
        k_lbvh_keys<<<G, B, 0, stream>>>(dTriRaw, n, l3, inv, keysA);
        if (!ok(cub::DeviceRadixSort::SortKeys(nullptr, tmpBytes, keysA, keysB, n, 0, 64, stream)) || !ok(cudaMalloc(&tmp, tmpBytes ? tmpBytes : 1))) break;
        if (!ok(cub::DeviceRadixSort::SortKeys(tmp, tmpBytes, keysA, keysB, n, 0, 64, stream))) break;
        ok(cudaMemsetAsync(visits, 0, (size_t)(n - 1) * 4, stream)); ok(cudaMemsetAsync(dDepth, 0, 4, stream));
        k_lbvh_internal<<<G, B, 0, stream>>>(keysB, n, children, parentI, parentL);
        k_lbvh_refit<<<G, B, 0, stream>>>(keysB, dTriRaw, n, children, parentI, parentL, visits, boxes);
        k_lbvh_emit<<<G, B, 0, stream>>>(keysB, dTriRaw, n, children, boxes, nodes);
        k_lbvh_depth<<<G, B, 0, stream>>>(n, parentI, parentL, dDepth);
        int depth = 0;
        if (!ok(cudaMemcpyAsync(&depth, dDepth, 4, cudaMemcpyDeviceToHost, stream)) || !ok(cudaStreamSynchronize(stream)) || !ok(cudaGetLastError())) break;
        if (outDepth) *outDepth = depth;
        if (depth + 1 > 46) { err = "linear BVH deeper than the ray caster's stack allows"; break; }
        *outNodes = nodes; *outNodeCount = 2 * n - 1; nodes = nullptr;
    } while (0);
    cudaFree(keysA); cudaFree(keysB); cudaFree(children); cudaFree(parentI); cudaFree(parentL); cudaFree(visits); cudaFree(boxes); cudaFree(dDepth); cudaFree(tmp); cudaFree(nodes);
    return err;
}

} // namespace pd
