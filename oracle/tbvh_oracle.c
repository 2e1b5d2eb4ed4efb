/* oracle/tbvh_oracle.c - plain-C CPU restatement of the tinybvh hot path.  See tbvh_oracle.h.
 *
 * TEST INFRASTRUCTURE ONLY - never linked or loaded by the product.
 *
 * Floating point.  The reference is built with gcc -O3 (CMakeLists.txt:49) and gcc's default
 * -ffp-contract=fast, so gcc fuses multiply-add pairs ("expect fma", tiny_bvh.h:3203).  Which pairs it
 * fuses is a compiler decision, so this file is compiled with -ffp-contract=off and spells every fused
 * operation as an explicit fmaf(), in the pairing read off the disassembly of the frozen oracle build
 * (g++ 13.3 -O3 -mavx2 -mfma, oracle/Makefile) and confirmed bit-for-bit by tests/test_oracle_pin.py:
 *   slab term           t = fmaf( bound, rD, -(O*rD) )                     (vfmsub, :3203-3214)
 *   h = cross(D,e2)     h.x = fmaf( D.y, e2.z, -(D.z*e2.y) )  (first product fused, second rounded)
 *   q = cross(s,e1)     q.x = fmaf( -s.z, e1.y, s.y*e1.z )    (second product fused, first rounded)
 *   a = dot(e1,h), dot(s,h), dot(e2,q)  = fmaf( z, z', fmaf( x, x', y*y' ) )
 *   dot(D,q)                            = fmaf( z, z', fmaf( y, y', x*x' ) )
 *   half area / SA      fmaf( ez, ex, fmaf( ey, ex, ey*ez ) )              (:460, :8477)
 *   bin index           trunc( fmaf( bmin+bmax, 0.5, -nmin ) * rpd )       (:2362, :2418)
 *   split cost          fmaf( c_int*rSAV, splitCost, c_trav )              (:2406)
 * The CUDA kernels use the same pairing through __fmaf_rn / __fmul_rn.
 */
#include "tbvh_oracle.h"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <stdio.h>
#include <pthread.h>
#include <unistd.h>

#define BVH_FAR 1e30f
#define BVHBINS 8

static inline float fminf_( float a, float b ) { return a < b ? a : b; }   /* tinybvh_min :445 */
static inline float fmaxf_( float a, float b ) { return a > b ? a : b; }   /* tinybvh_max :446 */

/* BVHBase::SA :8477 / tinybvh_half_area :460 (without its v.x < -BVH_FAR guard) */
static inline float half_area( float ex, float ey, float ez ) { return fmaf( ez, ex, fmaf( ey, ex, ey * ez ) ); }

/* ------------------------------------------------------------------------------------------------ build */

typedef struct { float bmin[3]; uint32_t primIdx; float bmax[3]; uint32_t clipped; } orc_fragment; /* BVHBase::Fragment :796-806 */

static inline int bin_of( float bmin, float bmax, float nmin, float rpd )
{
	/* :2362 - (int)(((bmin+bmax)*0.5f - nmin) * rpd); x86 cvttss2si semantics for NaN/overflow (-> INT_MIN) */
	float f = fmaf( bmin + bmax, 0.5f, -nmin ) * rpd;
	int bi = (f != f || f >= 2147483648.0f || f < -2147483648.0f) ? (int)0x80000000 : (int)f;
	return bi > 0 ? (bi < BVHBINS - 1 ? bi : BVHBINS - 1) : 0;  /* tinybvh_clamp :458 */
}

/* BuildAVXBinTask :6500-6502: cvtps2dq( fma( (bmax+bmin) - 2*nmin, rpd, -0.5 ) ), clamped to [0,7]; cvtps2dq rounds to nearest even
 * and yields INT_MIN for NaN / out of range */
static inline int bin_of_avx( float bmin, float bmax, float nmin2, float rpd )
{
	const float f = fmaf( (bmax + bmin) - nmin2, rpd, -0.5f );
	int bi = (f != f || f >= 2147483648.0f || f < -2147483648.0f) ? (int)0x80000000 : (int)lrintf( f );
	return bi > 0 ? (bi < BVHBINS - 1 ? bi : BVHBINS - 1) : 0;
}
/* BuildAVXSubtree :6629: (uint32_t)((bmax + bmin - nmin) * rpd) - 64-bit cvttss2si, low 32 bits, no clamp */
static inline uint32_t part_bin_avx( float bmin, float bmax, float nmin2, float rpd )
{
	const float f = (bmax + bmin - nmin2) * rpd;
	const long long v = (f != f || f >= 9223372036854775808.0f || f < -9223372036854775808.0f) ? (long long)0x8000000000000000ull : (long long)f;
	return (uint32_t)v;
}

/* flavour 0: BVH::Build (:2332-2461).  flavour 1: BVH::BuildAVX (:6400-6671) - the builder BuildDefault (:1817-1832), and with it
 * every derived layout's Build(), runs on x86.  Same algorithm; it differs in: the bin index (round-to-nearest of
 * fma(bmax+bmin-2*nmin, 3.99992/extent, -0.5), rpd = 0 on a zero extent); the partition's own bin function (truncation,
 * no -0.5, no clamp) which also decides the child counts; minDim = 1e-7 * root extent; planes tried in the order
 * 3,2,4,5,1,0,6 with both sides non-empty; cost = fma(lN, areaL, rN*areaR). */
static uint32_t build_impl( const float* verts, uint32_t primCount, orc_node* nodes, uint32_t* primIdx, float c_trav, float c_int, int flavour )
{
	orc_fragment* fragment = (orc_fragment*)malloc( (size_t)primCount * sizeof( orc_fragment ) );
	memset( &nodes[1], 0, sizeof( orc_node ) ); /* node 1 unused, :2285 */
	orc_node* root = &nodes[0];
	root->leftFirst = 0, root->triCount = primCount;
	float rmin[3] = { BVH_FAR, BVH_FAR, BVH_FAR }, rmax[3] = { -BVH_FAR, -BVH_FAR, -BVH_FAR };
	for (uint32_t i = 0; i < primCount; i++) /* PrepareBuild :2300-2308 */
	{
		const float* v0 = verts + (size_t)i * 12, * v1 = v0 + 4, * v2 = v0 + 8;
		for (int a = 0; a < 3; a++)
		{
			fragment[i].bmin[a] = fminf_( v0[a], fminf_( v1[a], v2[a] ) );
			fragment[i].bmax[a] = fmaxf_( v0[a], fmaxf_( v1[a], v2[a] ) );
			rmin[a] = fminf_( rmin[a], fragment[i].bmin[a] ), rmax[a] = fmaxf_( rmax[a], fragment[i].bmax[a] );
		}
		fragment[i].primIdx = i, fragment[i].clipped = 0, primIdx[i] = i;
	}
	root->minx = rmin[0], root->miny = rmin[1], root->minz = rmin[2];
	root->maxx = rmax[0], root->maxy = rmax[1], root->maxz = rmax[2];
	uint32_t newNodePtr = 2;
	/* Build(nodeIdx, depth) :2332-2461, non-threaded numbering */
	uint32_t task[256], taskCount = 0, nodeIdx = 0;
	const float mdf = flavour ? 1e-7f : 1e-20f;
	float minDim[3] = { (rmax[0] - rmin[0]) * mdf, (rmax[1] - rmin[1]) * mdf, (rmax[2] - rmin[2]) * mdf };
	float bestLMin[3] = { 0 }, bestLMax[3] = { 0 }, bestRMin[3] = { 0 }, bestRMax[3] = { 0 };
	while (1)
	{
		while (1)
		{
			orc_node* node = &nodes[nodeIdx];
			const float nmin3[3] = { node->minx, node->miny, node->minz }, nmax3[3] = { node->maxx, node->maxy, node->maxz };
			float binMin[3][BVHBINS][3], binMax[3][BVHBINS][3];
			uint32_t count[3][BVHBINS];
			for (int a = 0; a < 3; a++) for (int i = 0; i < BVHBINS; i++) for (int k = 0; k < 3; k++)
				binMin[a][i][k] = BVH_FAR, binMax[a][i][k] = -BVH_FAR;
			memset( count, 0, sizeof( count ) );
			float rpd3[3], nmin2[3] = { 0, 0, 0 };
			for (int a = 0; a < 3; a++)
			{
				if (!flavour) { rpd3[a] = (float)BVHBINS / (nmax3[a] - nmin3[a]); continue; }
				const float d = nmax3[a] - nmin3[a]; /* :6557-6559 */
				nmin2[a] = nmin3[a] * 2.0f, rpd3[a] = d == 0 ? 0.0f : (BVHBINS * 0.49999f) / d;
			}
			for (uint32_t i = 0; i < node->triCount; i++) /* binning :2357-2376 */
			{
				const orc_fragment* f = &fragment[primIdx[node->leftFirst + i]];
				for (int a = 0; a < 3; a++)
				{
					const int bi = flavour ? bin_of_avx( f->bmin[a], f->bmax[a], nmin2[a], rpd3[a] ) : bin_of( f->bmin[a], f->bmax[a], nmin3[a], rpd3[a] );
					for (int k = 0; k < 3; k++)
						binMin[a][bi][k] = fminf_( binMin[a][bi][k], f->bmin[k] ),
						binMax[a][bi][k] = fmaxf_( binMax[a][bi][k], f->bmax[k] );
					count[a][bi]++;
				}
			}
			/* sweep :2380-2405 */
			float splitCost = BVH_FAR;
			const float rSAV = 1.0f / half_area( nmax3[0] - nmin3[0], nmax3[1] - nmin3[1], nmax3[2] - nmin3[2] );
			uint32_t bestAxis = 0, bestPos = 0;
			for (int a = 0; a < 3; a++) if ((nmax3[a] - nmin3[a]) > minDim[a])
			{
				if (flavour)
				{
					/* :6597-6621 - prefix / suffix unions, then the planes in the order 3,2,4,5,1,0,6 */
					static const int order[7] = { 3, 2, 4, 5, 1, 0, 6 };
					float lmn[7][3], lmx[7][3], rmn[7][3], rmx[7][3];
					uint32_t lNs[7], rNs[7];
					float a1[3] = { BVH_FAR, BVH_FAR, BVH_FAR }, a2[3] = { -BVH_FAR, -BVH_FAR, -BVH_FAR };
					float b1[3] = { BVH_FAR, BVH_FAR, BVH_FAR }, b2[3] = { -BVH_FAR, -BVH_FAR, -BVH_FAR };
					uint32_t ln = 0, rn = 0;
					for (int i = 0; i < 7; i++)
					{
						for (int k = 0; k < 3; k++)
						{
							lmn[i][k] = a1[k] = fminf_( a1[k], binMin[a][i][k] ), lmx[i][k] = a2[k] = fmaxf_( a2[k], binMax[a][i][k] );
							rmn[6 - i][k] = b1[k] = fminf_( b1[k], binMin[a][7 - i][k] ), rmx[6 - i][k] = b2[k] = fmaxf_( b2[k], binMax[a][7 - i][k] );
						}
						ln += count[a][i], rn += count[a][7 - i];
						lNs[i] = ln, rNs[6 - i] = rn;
					}
					for (int o = 0; o < 7; o++)
					{
						const int i = order[o];
						if (lNs[i] == 0 || rNs[i] == 0) continue; /* PROCESS_PLANE: lN * rN != 0 */
						const float aL = half_area( lmx[i][0] - lmn[i][0], lmx[i][1] - lmn[i][1], lmx[i][2] - lmn[i][2] );
						const float aR = half_area( rmx[i][0] - rmn[i][0], rmx[i][1] - rmn[i][1], rmx[i][2] - rmn[i][2] );
						const float C = fmaf( (float)lNs[i], aL, aR * (float)rNs[i] );
						if (C < splitCost)
						{
							splitCost = C, bestAxis = a, bestPos = i;
							for (int k = 0; k < 3; k++) bestLMin[k] = lmn[i][k], bestLMax[k] = lmx[i][k], bestRMin[k] = rmn[i][k], bestRMax[k] = rmx[i][k];
						}
					}
					continue;
				}
				float lBMin[BVHBINS - 1][3], rBMin[BVHBINS - 1][3], lBMax[BVHBINS - 1][3], rBMax[BVHBINS - 1][3];
				float l1[3] = { BVH_FAR, BVH_FAR, BVH_FAR }, l2[3] = { -BVH_FAR, -BVH_FAR, -BVH_FAR };
				float r1[3] = { BVH_FAR, BVH_FAR, BVH_FAR }, r2[3] = { -BVH_FAR, -BVH_FAR, -BVH_FAR };
				float ANL[BVHBINS - 1], ANR[BVHBINS - 1];
				uint32_t lN = 0, rN = 0;
				for (int i = 0; i < BVHBINS - 1; i++)
				{
					for (int k = 0; k < 3; k++)
					{
						lBMin[i][k] = l1[k] = fminf_( l1[k], binMin[a][i][k] );
						rBMin[BVHBINS - 2 - i][k] = r1[k] = fminf_( r1[k], binMin[a][BVHBINS - 1 - i][k] );
						lBMax[i][k] = l2[k] = fmaxf_( l2[k], binMax[a][i][k] );
						rBMax[BVHBINS - 2 - i][k] = r2[k] = fmaxf_( r2[k], binMax[a][BVHBINS - 1 - i][k] );
					}
					lN += count[a][i], rN += count[a][BVHBINS - 1 - i];
					/* tinybvh_half_area :460 returns 0 when v.x < -BVH_FAR (cannot happen once lN > 0) */
					{
						const float ex = l2[0] - l1[0], ey = l2[1] - l1[1], ez = l2[2] - l1[2];
						ANL[i] = lN == 0 ? BVH_FAR : ((ex < -BVH_FAR ? 0 : half_area( ex, ey, ez )) * (float)lN);
					}
					{
						const float ex = r2[0] - r1[0], ey = r2[1] - r1[1], ez = r2[2] - r1[2];
						ANR[BVHBINS - 2 - i] = rN == 0 ? BVH_FAR : ((ex < -BVH_FAR ? 0 : half_area( ex, ey, ez )) * (float)rN);
					}
				}
				for (int i = 0; i < BVHBINS - 1; i++)
				{
					const float C = ANL[i] + ANR[i];
					if (C < splitCost)
					{
						splitCost = C, bestAxis = a, bestPos = i;
						for (int k = 0; k < 3; k++)
							bestLMin[k] = lBMin[i][k], bestRMin[k] = rBMin[i][k], bestLMax[k] = lBMax[i][k], bestRMax[k] = rBMax[i][k];
					}
				}
			}
			splitCost = fmaf( c_int * rSAV, splitCost, c_trav ); /* :2406 */
			const float noSplitCost = (float)node->triCount * c_int;
			if (splitCost >= noSplitCost) break;
			/* in-place partition :2414-2422 */
			uint32_t j = node->leftFirst + node->triCount, src = node->leftFirst;
			const float rpd = rpd3[bestAxis], nmin = nmin3[bestAxis];
			for (uint32_t i = 0; i < node->triCount; i++)
			{
				const orc_fragment* f = &fragment[primIdx[src]];
				const uint32_t bi = flavour ? part_bin_avx( f->bmin[bestAxis], f->bmax[bestAxis], nmin2[bestAxis], rpd ) : (uint32_t)bin_of( f->bmin[bestAxis], f->bmax[bestAxis], nmin, rpd );
				if (bi <= bestPos) src++;
				else { uint32_t t = primIdx[src]; primIdx[src] = primIdx[--j], primIdx[j] = t; }
			}
			const uint32_t leftCount = src - node->leftFirst, rightCount = node->triCount - leftCount;
			if (leftCount == 0 || rightCount == 0 || taskCount == 256) break;
			const uint32_t n = newNodePtr;
			newNodePtr += 2;
			nodes[n].minx = bestLMin[0], nodes[n].miny = bestLMin[1], nodes[n].minz = bestLMin[2];
			nodes[n].maxx = bestLMax[0], nodes[n].maxy = bestLMax[1], nodes[n].maxz = bestLMax[2];
			nodes[n].leftFirst = node->leftFirst, nodes[n].triCount = leftCount;
			nodes[n + 1].minx = bestRMin[0], nodes[n + 1].miny = bestRMin[1], nodes[n + 1].minz = bestRMin[2];
			nodes[n + 1].maxx = bestRMax[0], nodes[n + 1].maxy = bestRMax[1], nodes[n + 1].maxz = bestRMax[2];
			nodes[n + 1].leftFirst = j, nodes[n + 1].triCount = rightCount;
			node->leftFirst = n, node->triCount = 0;
			task[taskCount++] = n + 1, nodeIdx = n;
		}
		if (taskCount == 0) break; else nodeIdx = task[--taskCount];
	}
	free( fragment );
	return newNodePtr;
}

uint32_t orc_build( const float* verts, uint32_t primCount, orc_node* nodes, uint32_t* primIdx, float c_trav, float c_int )
{
	return build_impl( verts, primCount, nodes, primIdx, c_trav, c_int, 0 );
}
uint32_t orc_build_avx( const float* verts, uint32_t primCount, orc_node* nodes, uint32_t* primIdx, float c_trav, float c_int )
{
	return build_impl( verts, primCount, nodes, primIdx, c_trav, c_int, 1 );
}

/* -------------------------------------------------------------------------------------------- traversal */

typedef struct { float O[3]; uint32_t mask; float D[3]; uint32_t instIdx; float rD[3]; uint32_t pad; float t, u, v; uint32_t prim; uint8_t aux[64]; } orc_ray; /* Ray :688-709 */

/* MOLLER_TRUMBORE_TEST :1644-1656, IntersectTri :8508-8511 */
int orc_tri_test( const float* O, const float* D, const float* v0, const float* v1, const float* v2, float tmax, float* to, float* uo, float* vo )
{
	const float e1x = v1[0] - v0[0], e1y = v1[1] - v0[1], e1z = v1[2] - v0[2];
	const float e2x = v2[0] - v0[0], e2y = v2[1] - v0[1], e2z = v2[2] - v0[2];
	const float hx = fmaf( D[1], e2z, -(D[2] * e2y) );
	const float hy = fmaf( D[2], e2x, -(D[0] * e2z) );
	const float hz = fmaf( D[0], e2y, -(D[1] * e2x) );
	const float a = fmaf( e1z, hz, fmaf( e1x, hx, e1y * hy ) );
	if (fabsf( a ) < 0.000001f) return 0;
	const float f = 1 / a;
	const float sx = O[0] - v0[0], sy = O[1] - v0[1], sz = O[2] - v0[2];
	const float u = f * fmaf( hz, sz, fmaf( hx, sx, hy * sy ) );
	const float qx = fmaf( -e1y, sz, e1z * sy );
	const float qy = fmaf( -e1z, sx, e1x * sz );
	const float qz = fmaf( -e1x, sy, e1y * sx );
	const float v = f * fmaf( D[2], qz, fmaf( D[1], qy, D[0] * qx ) );
	if (u < 0 || v < 0 || u + v > 1) return 0;
	const float t = f * fmaf( e2z, qz, fmaf( e2x, qx, e2y * qy ) );
	if (t < 0 || t > tmax) return 0;
	*to = t, *uo = u, *vo = v;
	return 1;
}

/* SLAB_TEST_TWO_NODES :3202-3220 for one child; returns tmin or BVH_FAR */
static inline float slab( const orc_node* c, const float* rD, const float* ro, const int* pos, float tmax )
{
	const float* lo = &c->minx, * hi = &c->maxx;
	const float tx1 = fmaf( pos[0] ? lo[0] : hi[0], rD[0], -ro[0] ), tx2 = fmaf( pos[0] ? hi[0] : lo[0], rD[0], -ro[0] );
	const float ty1 = fmaf( pos[1] ? lo[1] : hi[1], rD[1], -ro[1] ), ty2 = fmaf( pos[1] ? hi[1] : lo[1], rD[1], -ro[1] );
	const float tz1 = fmaf( pos[2] ? lo[2] : hi[2], rD[2], -ro[2] ), tz2 = fmaf( pos[2] ? hi[2] : lo[2], rD[2], -ro[2] );
	const float tmin = fmaxf_( fmaxf_( tx1, ty1 ), fmaxf_( tz1, 0.0f ) );
	const float tmx = fminf_( fminf_( tx2, ty2 ), fminf_( tz2, tmax ) );
	return tmx >= tmin ? tmin : BVH_FAR;
}

static int intersect1( const orc_node* nodes, const uint32_t* primIdx, const float* verts, orc_ray* ray, int anyhit )
{
	const orc_node* node = &nodes[0], * stack[256];
	uint32_t stackPtr = 0;
	const int pos[3] = { ray->D[0] >= 0, ray->D[1] >= 0, ray->D[2] >= 0 };
	const float ro[3] = { ray->O[0] * ray->rD[0], ray->O[1] * ray->rD[1], ray->O[2] * ray->rD[2] };
	while (1)
	{
		if (node->triCount > 0)
		{
			for (uint32_t i = 0; i < node->triCount; i++)
			{
				const uint32_t pi = primIdx[node->leftFirst + i];
				const float* v0 = verts + (size_t)pi * 12;
				float t, u, v;
				if (orc_tri_test( ray->O, ray->D, v0, v0 + 4, v0 + 8, ray->t, &t, &u, &v ))
				{
					if (anyhit) return 1;
					ray->t = t, ray->u = u, ray->v = v, ray->prim = pi;
					ray->pad = ray->instIdx; /* hit.inst = ray.instIdx (IntersectTri :8525, INST_IDX_BITS == 32): byte 44 of the record */
				}
			}
			if (stackPtr == 0) break; else node = stack[--stackPtr];
			continue;
		}
		const orc_node* child1 = &nodes[node->leftFirst], * child2 = &nodes[node->leftFirst + 1];
		float dist1 = slab( child1, ray->rD, ro, pos, ray->t ), dist2 = slab( child2, ray->rD, ro, pos, ray->t );
		if (dist1 > dist2) { float t = dist1; dist1 = dist2, dist2 = t; const orc_node* c = child1; child1 = child2, child2 = c; }
		if (dist1 == BVH_FAR) { if (stackPtr == 0) break; else node = stack[--stackPtr]; }
		else { node = child1; if (dist2 != BVH_FAR) stack[stackPtr++] = child2; }
	}
	return 0;
}

/* worker pool: chunks of 4096 rays (a multiple of 32, so occlusion words never straddle) off an atomic counter */
typedef struct { const orc_node* nodes; const uint32_t* primIdx; const float* verts; orc_ray* rays; uint64_t n; uint32_t* bits; uint64_t next; } orc_job;
#define ORC_CHUNK 4096
static void* orc_worker( void* arg )
{
	orc_job* j = (orc_job*)arg;
	for (;;)
	{
		const uint64_t s = __atomic_fetch_add( &j->next, ORC_CHUNK, __ATOMIC_RELAXED );
		if (s >= j->n) break;
		const uint64_t e = s + ORC_CHUNK < j->n ? s + ORC_CHUNK : j->n;
		if (!j->bits) for (uint64_t i = s; i < e; i++) intersect1( j->nodes, j->primIdx, j->verts, &j->rays[i], 0 );
		else for (uint64_t w = s / 32; w * 32 < e; w++)
		{
			uint32_t m = 0;
			for (int b = 0; b < 32 && w * 32 + b < e; b++)
			{
				orc_ray tmp = j->rays[w * 32 + b];
				if (intersect1( j->nodes, j->primIdx, j->verts, &tmp, 1 )) m |= 1u << b;
			}
			j->bits[w] = m;
		}
	}
	return 0;
}
static void orc_run( orc_job* j, int threads )
{
	if (threads <= 0) threads = (int)sysconf( _SC_NPROCESSORS_ONLN );
	if (threads > 512) threads = 512;
	if (threads <= 1 || j->n <= ORC_CHUNK) { orc_worker( j ); return; }
	pthread_t tid[512];
	for (int t = 0; t < threads; t++) pthread_create( &tid[t], 0, orc_worker, j );
	for (int t = 0; t < threads; t++) pthread_join( tid[t], 0 );
}

void orc_intersect( const orc_node* nodes, const uint32_t* primIdx, const float* verts, void* rays, uint64_t n, int threads )
{
	orc_job j = { nodes, primIdx, verts, (orc_ray*)rays, n, 0, 0 };
	orc_run( &j, threads );
}

void orc_occluded( const orc_node* nodes, const uint32_t* primIdx, const float* verts, const void* rays, uint64_t n, uint32_t* bits, int threads )
{
	orc_job j = { nodes, primIdx, verts, (orc_ray*)rays, n, bits, 0 };
	orc_run( &j, threads );
}

/* BVH_GPU::ConvertFrom :4612-4655 */
/* -------------------------------------------------------------------------------------------- TLAS / BLAS
 * BVH::IntersectTLAS :3306-3380 and IsOccludedTLAS :3455-3519 over BVH-layout BLASses (INST_IDX_BITS == 32: the instance
 * number travels in hit.inst, byte 44 of the Ray record).  Per instance of a TLAS leaf: skip unless inst.mask & ray.mask,
 * O' = transform_point( O, invTransform ), D' = transform_vector( D, invTransform ) (:513-527; the frozen build computes a
 * row as fma( Tz, z, fma( Tx, x, Ty*y ) ) + Tw and divides by w only when w != 1), rD' = safercp( D' ) (:442), then the
 * BLAS is traversed by BVH::Intersect / IsOccluded with the running hit. */
static inline float safercp( float x ) { if (x > 1e-12f || x < -1e-12f) return 1.0f / x; else return x >= 0 ? BVH_FAR : -BVH_FAR; }
static void transform_ray( const float* T, const orc_ray* ray, orc_ray* temp )
{
	const float* O = ray->O, * D = ray->D;
	float r[3];
	for (int k = 0; k < 3; k++) r[k] = fmaf( T[k * 4 + 2], O[2], fmaf( T[k * 4], O[0], T[k * 4 + 1] * O[1] ) ) + T[k * 4 + 3];
	const float w = fmaf( O[2], T[14], fmaf( O[0], T[12], O[1] * T[13] ) ) + T[15];
	if (!(w == 1)) { const float rw = 1.0f / w; for (int k = 0; k < 3; k++) r[k] = r[k] * rw; }
	for (int k = 0; k < 3; k++)
	{
		temp->O[k] = r[k];
		temp->D[k] = fmaf( T[k * 4 + 2], D[2], fmaf( T[k * 4], D[0], T[k * 4 + 1] * D[1] ) );
		temp->rD[k] = safercp( temp->D[k] );
	}
}
/* the BLAS step of one TLAS-leaf instance: walks `temp` (the transformed ray, hit copied from the world-space ray) through BLAS
 * blasIdx and leaves the updated hit in it; returns 1 when an any-hit query found an occluder */
static int walk_bvh_blas( const void* user, uint32_t blasIdx, void* temp, int anyhit )
{
	const orc_blas* b = (const orc_blas*)user + blasIdx;
	return intersect1( b->nodes, b->primIdx, b->verts, (orc_ray*)temp, anyhit );
}
int orc_tlas_walk1( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, void* ray_, int anyhit, orc_blas_walk walk, const void* user )
{
	orc_ray* ray = (orc_ray*)ray_;
	const orc_node* node = &nodes[0], * stack[64];
	uint32_t stackPtr = 0;
	const int pos[3] = { ray->D[0] >= 0, ray->D[1] >= 0, ray->D[2] >= 0 };
	const float ro[3] = { ray->O[0] * ray->rD[0], ray->O[1] * ray->rD[1], ray->O[2] * ray->rD[2] };
	while (1)
	{
		if (node->triCount > 0)
		{
			for (uint32_t i = 0; i < node->triCount; i++)
			{
				const uint32_t instIdx = primIdx[node->leftFirst + i];
				const orc_instance* in = &inst[instIdx];
				if (!(in->mask & ray->mask)) continue;
				orc_ray temp;
				memset( &temp, 0, sizeof( temp ) );
				temp.mask = 0xFFFF; /* Ray() = default: mask = RAY_MASK_INTERSECT_ALL */
				transform_ray( in->invTransform, ray, &temp );
				temp.instIdx = instIdx; /* << (32 - INST_IDX_BITS) = << 0 */
				temp.pad = ray->pad, temp.t = ray->t, temp.u = ray->u, temp.v = ray->v, temp.prim = ray->prim; /* temp.hit = ray.hit */
				if (walk( user, in->blasIdx, &temp, anyhit )) return 1;
				if (!anyhit) ray->pad = temp.pad, ray->t = temp.t, ray->u = temp.u, ray->v = temp.v, ray->prim = temp.prim; /* ray.hit = temp.hit */
			}
			if (stackPtr == 0) break; else node = stack[--stackPtr];
			continue;
		}
		const orc_node* child1 = &nodes[node->leftFirst], * child2 = &nodes[node->leftFirst + 1];
		float dist1 = slab( child1, ray->rD, ro, pos, ray->t ), dist2 = slab( child2, ray->rD, ro, pos, ray->t );
		if (dist1 > dist2) { float t = dist1; dist1 = dist2, dist2 = t; const orc_node* c = child1; child1 = child2, child2 = c; }
		if (dist1 == BVH_FAR) { if (stackPtr == 0) break; else node = stack[--stackPtr]; }
		else { node = child1; if (dist2 != BVH_FAR) stack[stackPtr++] = child2; }
	}
	return 0;
}
static int intersect_tlas1( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_blas* blas, orc_ray* ray, int anyhit )
{
	return orc_tlas_walk1( nodes, primIdx, inst, ray, anyhit, walk_bvh_blas, blas );
}
void orc_intersect_tlas( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_blas* blas, void* rays, uint64_t n )
{
	for (uint64_t i = 0; i < n; i++) intersect_tlas1( nodes, primIdx, inst, blas, (orc_ray*)rays + i, 0 );
}
void orc_occluded_tlas( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_blas* blas, const void* rays, uint64_t n, uint32_t* bits )
{
	for (uint64_t i = 0; i < n; i++)
	{
		orc_ray tmp = ((const orc_ray*)rays)[i];
		if (intersect_tlas1( nodes, primIdx, inst, blas, &tmp, 1 )) bits[i >> 5] |= 1u << (i & 31);
	}
}

#define UPD_ATTR

/* BLASInstance::InvertTransform (:8402-8428) + Update (:8386-8399).  The frozen reference build vectorises the sixteen cofactor
 * sums, and the lanes of different rows end up with different fused multiply-adds; the shapes below were identified by
 * exhaustive search over every contraction gcc may legally form (3 x 2^4 per cofactor, 12 for the determinant, 6^4 for the
 * corner transform) against BLASInstance::Update on 2,000 random matrices, and are pinned by the tests.  A cofactor is
 * s0*(a0*b0)*c0 + ... + s5*(a5*b5)*c5 with p_k = round( a_k*b_k ):
 *   rows 0,1 : t0 = round( p0*c0 ); acc = fma( +-p1, c1, t0 ); then fma( +-p_k, c_k, acc ) for k = 2..5
 *   row 2    : acc = fma( p0, c0, +-round( p1*c1 ) ); k=2: acc +- round( p2*c2 ); k=3,4: fma; k=5: acc +- round( p5*c5 )
 *   row 3    : acc = fma( p0, c0, +-round( p1*c1 ) ); k=2: fma; k=3,4: acc +- round(..); k=5: fma
 *   det      = fma( T3, c12, fma( T2, c8, fma( T0, c0, round( T1*c4 ) ) ) ); every cell is then multiplied by 1/det
 *   corner   : row = fma( Tz, z, fma( Tx, x, round( Ty*y ) ) ) + Tw, divided by w only when w != 1 */
typedef struct { signed char s; unsigned char a, b, c; } UPD_TERM;
static const UPD_TERM UPD_E[16][6] = {
 {{+1,5,10,15},{-1,5,11,14},{-1,9,6,15},{+1,9,7,14},{+1,13,6,11},{-1,13,7,10}}, {{-1,1,10,15},{+1,1,11,14},{+1,9,2,15},{-1,9,3,14},{-1,13,2,11},{+1,13,3,10}},
 {{+1,1,6,15},{-1,1,7,14},{-1,5,2,15},{+1,5,3,14},{+1,13,2,7},{-1,13,3,6}}, {{-1,1,6,11},{+1,1,7,10},{+1,5,2,11},{-1,5,3,10},{-1,9,2,7},{+1,9,3,6}},
 {{-1,4,10,15},{+1,4,11,14},{+1,8,6,15},{-1,8,7,14},{-1,12,6,11},{+1,12,7,10}}, {{+1,0,10,15},{-1,0,11,14},{-1,8,2,15},{+1,8,3,14},{+1,12,2,11},{-1,12,3,10}},
 {{-1,0,6,15},{+1,0,7,14},{+1,4,2,15},{-1,4,3,14},{-1,12,2,7},{+1,12,3,6}}, {{+1,0,6,11},{-1,0,7,10},{-1,4,2,11},{+1,4,3,10},{+1,8,2,7},{-1,8,3,6}},
 {{+1,4,9,15},{-1,4,11,13},{-1,8,5,15},{+1,8,7,13},{+1,12,5,11},{-1,12,7,9}}, {{-1,0,9,15},{+1,0,11,13},{+1,8,1,15},{-1,8,3,13},{-1,12,1,11},{+1,12,3,9}},
 {{+1,0,5,15},{-1,0,7,13},{-1,4,1,15},{+1,4,3,13},{+1,12,1,7},{-1,12,3,5}}, {{-1,0,5,11},{+1,0,7,9},{+1,4,1,11},{-1,4,3,9},{-1,8,1,7},{+1,8,3,5}},
 {{-1,4,9,14},{+1,4,10,13},{+1,8,5,14},{-1,8,6,13},{-1,12,5,10},{+1,12,6,9}}, {{+1,0,9,14},{-1,0,10,13},{-1,8,1,14},{+1,8,2,13},{+1,12,1,10},{-1,12,2,9}},
 {{-1,0,5,14},{+1,0,6,13},{+1,4,1,14},{-1,4,2,13},{-1,12,1,6},{+1,12,2,5}}, {{+1,0,5,10},{-1,0,6,9},{-1,4,1,10},{+1,4,2,9},{+1,8,1,6},{-1,8,2,5}} };
UPD_ATTR static float upd_cofactor( const float* T, const UPD_TERM* e, const int first_fused_left, const unsigned fused_mask )
{
	float p[6];
	for (int k = 0; k < 6; k++) p[k] = T[e[k].a] * T[e[k].b];
	const float p0 = e[0].s < 0 ? -p[0] : p[0];
	float acc;
	if (first_fused_left) { const float t1 = p[1] * T[e[1].c]; acc = fmaf( p0, T[e[0].c], e[1].s > 0 ? t1 : -t1 ); }
	else { const float t0 = p0 * T[e[0].c]; acc = fmaf( e[1].s > 0 ? p[1] : -p[1], T[e[1].c], t0 ); }
	for (int k = 2; k < 6; k++)
	{
		if (fused_mask & (1u << (k - 2))) acc = fmaf( e[k].s > 0 ? p[k] : -p[k], T[e[k].c], acc );
		else { const float t = p[k] * T[e[k].c]; acc = e[k].s > 0 ? acc + t : acc - t; }
	}
	return acc;
}
UPD_ATTR static void upd_instance( float* T /* transform, 16 */, float* iT /* invTransform, 16 */, float* aabbMin, float* aabbMax, const float* bmin, const float* bmax )
{
	float c[16];
	for (int k = 0; k < 16; k++) c[k] = k < 8 ? upd_cofactor( T, UPD_E[k], 0, 15u ) : k < 12 ? upd_cofactor( T, UPD_E[k], 1, 6u ) : upd_cofactor( T, UPD_E[k], 1, 9u );
	const float t14 = T[1] * c[4];
	const float det = fmaf( T[3], c[12], fmaf( T[2], c[8], fmaf( T[0], c[0], t14 ) ) );
	if (det == 0) { for (int k = 0; k < 16; k++) iT[k] = c[k]; } /* "invert failed": the reference returns with the cofactors stored */
	else { const float invdet = 1.0f / det; for (int k = 0; k < 16; k++) iT[k] = c[k] * invdet; }
	for (int k = 0; k < 3; k++) aabbMin[k] = 1e30f, aabbMax[k] = -1e30f;
	for (int j = 0; j < 8; j++)
	{
		const float p[3] = { j & 1 ? bmax[0] : bmin[0], j & 2 ? bmax[1] : bmin[1], j & 4 ? bmax[2] : bmin[2] };
		float r[3];
		for (int k = 0; k < 3; k++) { const float ty = T[k * 4 + 1] * p[1]; r[k] = fmaf( T[k * 4 + 2], p[2], fmaf( T[k * 4], p[0], ty ) ) + T[k * 4 + 3]; }
		const float wy = T[13] * p[1];
		const float w = fmaf( T[14], p[2], fmaf( T[12], p[0], wy ) ) + T[15];
		if (!(w == 1)) { const float rw = 1.0f / w; r[0] = r[0] * rw, r[1] = r[1] * rw, r[2] = r[2] * rw; }
		for (int k = 0; k < 3; k++) aabbMin[k] = aabbMin[k] < r[k] ? aabbMin[k] : r[k], aabbMax[k] = aabbMax[k] > r[k] ? aabbMax[k] : r[k];
	}
}
void orc_instance_update( orc_instance* inst, const float* bmin, const float* bmax ) { upd_instance( inst->transform, inst->invTransform, inst->aabbMin, inst->aabbMax, bmin, bmax ); }

uint32_t orc_to_bvh_gpu( const orc_node* nodes, orc_node_gpu* out )
{
	uint32_t newNodePtr = 0, nodeIdx = 0, stack[512], stackPtr = 0;
	while (1)
	{
		const orc_node* orig = &nodes[nodeIdx];
		const uint32_t idx = newNodePtr++;
		memset( &out[idx], 0, sizeof( orc_node_gpu ) );
		if (orig->triCount > 0)
		{
			out[idx].triCount = orig->triCount, out[idx].firstTri = orig->leftFirst;
			if (!stackPtr) break;
			nodeIdx = stack[--stackPtr];
			const uint32_t parent = stack[--stackPtr];
			out[parent].right = newNodePtr;
		}
		else
		{
			const orc_node* l = &nodes[orig->leftFirst], * r = &nodes[orig->leftFirst + 1];
			memcpy( out[idx].lmin, &l->minx, 12 ), memcpy( out[idx].lmax, &l->maxx, 12 );
			memcpy( out[idx].rmin, &r->minx, 12 ), memcpy( out[idx].rmax, &r->maxx, 12 );
			out[idx].left = newNodePtr;
			stack[stackPtr++] = idx, stack[stackPtr++] = orig->leftFirst + 1;
			nodeIdx = orig->leftFirst;
		}
	}
	return newNodePtr;
}

/* BVH::SAHCost :1889-1898 (quality metric; recursion as the reference) */
static float sah_rec( const orc_node* nodes, uint32_t i, float c_trav, float c_int )
{
	const orc_node* n = &nodes[i];
	const float sa = half_area( n->maxx - n->minx, n->maxy - n->miny, n->maxz - n->minz );
	if (n->triCount > 0) return c_int * sa * n->triCount;
	return c_trav * sa + sah_rec( nodes, n->leftFirst, c_trav, c_int ) + sah_rec( nodes, n->leftFirst + 1, c_trav, c_int );
}
/* BVH::Refit :3055-3093 (no vertIdx): nodes backwards, leaves from their triangles' current vertices, interior nodes from
 * their children; node 1 is skipped. */
void orc_refit( orc_node* nodes, uint32_t usedNodes, const uint32_t* primIdx, const float* verts )
{
	for (int32_t i = (int32_t)usedNodes - 1; i >= 0; i--) if (i != 1)
	{
		orc_node* n = &nodes[i];
		if (n->triCount)
		{
			float bmin[3] = { BVH_FAR, BVH_FAR, BVH_FAR }, bmax[3] = { -BVH_FAR, -BVH_FAR, -BVH_FAR };
			for (uint32_t j = 0; j < n->triCount; j++)
			{
				const float* v0 = verts + (size_t)primIdx[n->leftFirst + j] * 12, * v1 = v0 + 4, * v2 = v0 + 8;
				for (int a = 0; a < 3; a++)
				{
					const float t1 = v0[a] < bmin[a] ? v0[a] : bmin[a], t2 = v0[a] > bmax[a] ? v0[a] : bmax[a];
					const float t3 = v1[a] < v2[a] ? v1[a] : v2[a], t4 = v1[a] > v2[a] ? v1[a] : v2[a];
					bmin[a] = t1 < t3 ? t1 : t3, bmax[a] = t2 > t4 ? t2 : t4;
				}
			}
			n->minx = bmin[0], n->miny = bmin[1], n->minz = bmin[2], n->maxx = bmax[0], n->maxy = bmax[1], n->maxz = bmax[2];
			continue;
		}
		const orc_node* l = &nodes[n->leftFirst], * r = l + 1;
		n->minx = l->minx < r->minx ? l->minx : r->minx, n->miny = l->miny < r->miny ? l->miny : r->miny, n->minz = l->minz < r->minz ? l->minz : r->minz;
		n->maxx = l->maxx > r->maxx ? l->maxx : r->maxx, n->maxy = l->maxy > r->maxy ? l->maxy : r->maxy, n->maxz = l->maxz > r->maxz ? l->maxz : r->maxz;
	}
}

float orc_sah_cost( const orc_node* nodes, uint32_t nodeIdx, float c_trav, float c_int )
{
	const orc_node* n = &nodes[nodeIdx];
	const float cost = sah_rec( nodes, nodeIdx, c_trav, c_int );
	return nodeIdx == 0 ? cost / half_area( n->maxx - n->minx, n->maxy - n->miny, n->maxz - n->minz ) : cost;
}
