/* oracle/tbvh_oracle.h - plain-C CPU restatement of the tinybvh hot path (the "port" oracle).
 *
 * TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference arms may load this library; the product (tinybvh_b200/) never does.
 *
 * Parity status: PINNED - tests/test_oracle_pin.py checks every function here bit-for-bit against the
 * unmodified reference compiled into oracle/_ref (oracle/ref_wrap.cpp) and against the committed golden
 * vectors under tests/golden/ that were produced by that build (tools/make_golden.py).
 *
 * All file:line citations are /root/reference/tiny_bvh.h.
 */
#ifndef TBVH_ORACLE_H
#define TBVH_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { float minx, miny, minz; uint32_t leftFirst; float maxx, maxy, maxz; uint32_t triCount; } orc_node;   /* BVH::BVHNode :861-869 */
typedef struct { float lmin[3]; uint32_t left; float lmax[3]; uint32_t right; float rmin[3]; uint32_t triCount; float rmax[3]; uint32_t firstTri; } orc_node_gpu; /* BVH_GPU::BVHNode :1095-1105 */

/* BVH::PrepareBuild (:2261) + BVH::Build(nodeIdx,depth) (:2332), single-threaded numbering.
 * verts: primCount*3 float4; nodes: room for 2*primCount; primIdx: room for primCount. Returns usedNodes. */
uint32_t orc_build( const float* verts, uint32_t primCount, orc_node* nodes, uint32_t* primIdx, float c_trav, float c_int );
/* BVH::PrepareAVXBuild (:6424) + BVH::BuildAVXSubtree (:6529) = what BuildDefault (:1817) runs on x86, single-threaded numbering */
uint32_t orc_build_avx( const float* verts, uint32_t primCount, orc_node* nodes, uint32_t* primIdx, float c_trav, float c_int );

/* BVH::BuildHQ (SBVH, :2623-3040) + Compact (:3733), single-threaded; tbvh_oracle_hq.c.  nodes: room for 3*primCount+2,
 * primIdx: room for primCount + primCount/2.  Returns usedNodes. */
uint32_t orc_build_hq( const float* verts, uint32_t primCount, orc_node* nodes, uint32_t* primIdx, uint32_t* idxCount, uint32_t* usedIdx, float c_trav, float c_int );

/* test hooks for the two geometric helpers of the SBVH build: BVH::ClipFrag (:8614) and BVH::SplitFrag (:8731) on one
 * 32-byte Fragment record {bmin[3], primIdx, bmax[3], clipped} (:792) */
int orc_clip_frag( const float* verts, const void* orig, void* out, const float* bmin, const float* bmax, const float* minDim, uint32_t axis );
void orc_split_frag( const float* verts, const void* orig, void* left, void* right, const float* minDim, uint32_t axis, float pos, int* leftOk, int* rightOk );

/* BVH::Intersect (:3222,:3247) / BVH::IsOccluded (:3382,:3407) over 128-byte host Ray records, in place.
 * threads<=0 -> all cores (OpenMP). */
void orc_intersect( const orc_node* nodes, const uint32_t* primIdx, const float* verts, void* rays, uint64_t n, int threads );
void orc_occluded( const orc_node* nodes, const uint32_t* primIdx, const float* verts, const void* rays, uint64_t n, uint32_t* bits, int threads );

/* TLAS / BLAS: BVH::IntersectTLAS (:3306) / IsOccludedTLAS (:3455) with BVH-layout BLASses and INST_IDX_BITS == 32 (hit.inst
 * at byte 44 of the Ray record).  inst: the reference's 192-byte BLASInstance records (:1443), already Update()d;
 * blas[k]: the k-th BLAS (its node array, primIdx and vertices).  bits must be zeroed by the caller. */
typedef struct { float transform[16], invTransform[16]; float aabbMin[3]; uint32_t blasIdx; float aabbMax[3]; uint32_t mask; uint32_t dummy[8]; } orc_instance;
typedef struct { const orc_node* nodes; const uint32_t* primIdx; const float* verts; } orc_blas;
/* BLASInstance::Update (:8386): inverse transform + world box of a BLAS whose root box is [bmin, bmax] */
void orc_instance_update( orc_instance* inst, const float* bmin, const float* bmax );
void orc_intersect_tlas( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_blas* blas, void* rays, uint64_t n );
void orc_occluded_tlas( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_blas* blas, const void* rays, uint64_t n, uint32_t* bits );
/* the TLAS walk of IntersectTLAS / IsOccludedTLAS for ONE ray with the per-instance BLAS step supplied by the caller: walk( user, blasIdx,
 * temp, anyhit ) receives the transformed ray (hit = the world-space ray's current hit, instIdx set) and leaves the updated hit in it;
 * a non-zero return ends an any-hit query */
typedef int (*orc_blas_walk)( const void* user, uint32_t blasIdx, void* temp_ray, int anyhit );
int orc_tlas_walk1( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, void* ray, int anyhit, orc_blas_walk walk, const void* user );

/* One Moeller-Trumbore test in the oracle's arithmetic (MOLLER_TRUMBORE_TEST :1644-1656).
 * Returns 1 and writes t,u,v when the triangle is accepted for a ray with the given tmax. */
int orc_tri_test( const float* O, const float* D, const float* v0, const float* v1, const float* v2, float tmax, float* t, float* u, float* v );

/* BVH_GPU::ConvertFrom (:4612): DFS re-layout into 64-byte Aila-Laine nodes. Returns usedNodes. */
uint32_t orc_to_bvh_gpu( const orc_node* nodes, orc_node_gpu* out );

/* BVH::Refit (:3055-3093), flat (non-indexed) geometry: boxes recomputed in place from verts (primCount*3 float4) */
void orc_refit( orc_node* nodes, uint32_t usedNodes, const uint32_t* primIdx, const float* verts );

/* The conversion chain of BVH8_CWBVH::Build / BuildHQ (:5822-5866) applied to a BVH2: SplitLeafs( 3 ) :1988, MBVH<8>::ConvertFrom
 * :4975, BVH8_CWBVH::ConvertFrom :5884 (tbvh_oracle_cwbvh.c).  data: triCount * 5 float4 blocks, tris: idxCount * 3 float4; both are
 * zeroed first, as the reference does.  Returns usedBlocks. */
uint32_t orc_cwbvh_from_bvh( const orc_node* nodes, uint32_t usedNodes, const uint32_t* primIdx, uint32_t idxCount, const float* verts, uint32_t triCount, float* data, float* tris );
/* BVH8_CWBVH::Intersect (:7046-7154), the reference's CPU walk of the compressed layout, over 128-byte Ray records in place */
void orc_cwbvh_intersect( const float* bvh8Data, const float* bvh8Tris, void* rays, uint64_t n );
/* A TLAS over BVH8_CWBVH BLASses - the arrangement of the reference's GPU path (traverse_tlas.cl:13-107: BVH2 TLAS, per-instance ray
 * transform, traverse_cwbvh per BLAS, `if (blasHit.x < hit.x) hit = blasHit` with the instance attached); the reference's CPU
 * IntersectTLAS does not accept LAYOUT_CWBVH BLASses (:3339), so this is a COMPOSITION of two pinned pieces, not a pinned function:
 * orc_tlas_walk1 (IntersectTLAS :3306, transform and safercp as on the CPU) with BVH8_CWBVH::Intersect (:7046) as the BLAS step; a
 * BLAS that finds nothing closer leaves the hit as it was; occlusion = a BLAS walk that ends below the ray's t (FALLBACK_SHADOW_QUERY). */
typedef struct { const float* bvh8Data; const float* bvh8Tris; } orc_cwblas;
void orc_intersect_tlas_cw( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_cwblas* blas, void* rays, uint64_t n );
void orc_occluded_tlas_cw( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_cwblas* blas, const void* rays, uint64_t n, uint32_t* bits );

/* BVH::SAHCost (:1889) */
float orc_sah_cost( const orc_node* nodes, uint32_t nodeIdx, float c_trav, float c_int );

#ifdef __cplusplus
}
#endif
#endif
