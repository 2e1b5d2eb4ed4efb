/* oracle/tbvh_oracle_cwbvh.c - plain-C restatement of the reference's CWBVH chain and of its CPU walk of that layout.
 *
 * TEST INFRASTRUCTURE ONLY (see tbvh_oracle.h).  Follows tiny_bvh.h: BVH::SplitLeafs :1988-2017, MBVH<8>::ConvertFrom :4975-5048,
 * BVH8_CWBVH::ConvertFrom :5884-6018 (the conversion part of BVH8_CWBVH::Build / BuildHQ :5822-5866) and
 * BVH8_CWBVH::Intersect :7046-7154.  Compiled with -ffp-contract=off; the fused multiply-adds of the frozen reference build
 * are explicit: slab terms  fma( q, 2^e * rD, -(O - p) * rD )  and the Moeller-Trumbore pairing of tbvh_oracle.c.
 * Pinned byte for byte (bvh8Data, bvh8Tris) and bit for bit (hits) against the compiled reference in tests/test_oracle_pin.py.
 */
#include "tbvh_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define BVH_FAR 1e30f

typedef struct { float mn[3]; uint32_t firstTri; float mx[3]; uint32_t triCount; uint32_t child[8]; uint32_t childCount; uint32_t dummy[3]; } mnode; /* MBVH<8>::MBVHNode :1213 */

static inline float sa3( const float* mn, const float* mx ) /* BVHBase::SA :8477 */
{
	const float ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
	return fmaf( ez, ex, fmaf( ey, ex, ey * ez ) );
}

/* nodes: the BVH2 (Build / BuildAVX / BuildHQ output) with room for usedNodes + usedNodes/2 + 2 records; returns the new usedNodes */
static uint32_t split_leafs( orc_node* n, uint32_t newNodePtr, const uint32_t maxPrims ) /* :1988 */
{
	uint32_t stack[64], stackPtr = 0, nodeIdx = 0;
	while (1)
	{
		orc_node* node = &n[nodeIdx];
		if (node->triCount > 0)
		{
			if (node->triCount > maxPrims)
			{
				orc_node* left = &n[newNodePtr], * right = &n[newNodePtr + 1];
				*left = *node, *right = *node;
				right->leftFirst = node->leftFirst + maxPrims, right->triCount = node->triCount - maxPrims;
				left->triCount = maxPrims, node->leftFirst = newNodePtr, node->triCount = 0, newNodePtr += 2;
			}
			else
			{
				if (!stackPtr) break;
				nodeIdx = stack[--stackPtr];
			}
		}
		else nodeIdx = node->leftFirst, stack[stackPtr++] = node->leftFirst + 1;
	}
	return newNodePtr;
}

static void mbvh8_from_bvh( const orc_node* b, const uint32_t usedNodes, mnode* m ) /* :4975 */
{
	for (uint32_t i = 0; i < usedNodes; i++) if (i != 1)
	{
		const orc_node* o = &b[i];
		mnode* nd = &m[i];
		nd->mn[0] = o->minx, nd->mn[1] = o->miny, nd->mn[2] = o->minz, nd->mx[0] = o->maxx, nd->mx[1] = o->maxy, nd->mx[2] = o->maxz;
		if (o->triCount > 0) nd->triCount = o->triCount, nd->firstTri = o->leftFirst;
		else nd->child[0] = o->leftFirst, nd->child[1] = o->leftFirst + 1, nd->childCount = 2;
	}
	uint32_t stack[128], stackPtr = 0, nodeIdx = 0;
	while (1)
	{
		mnode* node = &m[nodeIdx];
		while (node->childCount < 8)
		{
			int32_t bestChild = -1;
			float bestChildSA = 0;
			for (uint32_t i = 0; i < node->childCount; i++)
			{
				const mnode* child = &m[node->child[i]];
				if (child->triCount == 0 && node->childCount - 1 + child->childCount <= 8)
				{
					const float childSA = sa3( child->mn, child->mx );
					if (childSA > bestChildSA) bestChild = (int32_t)i, bestChildSA = childSA;
				}
			}
			if (bestChild == -1) break;
			const mnode* child = &m[node->child[bestChild]];
			node->child[bestChild] = child->child[0];
			for (uint32_t i = 1; i < child->childCount; i++) node->child[node->childCount++] = child->child[i];
		}
		for (uint32_t i = 0; i < node->childCount; i++)
		{
			const uint32_t childIdx = node->child[i];
			if (m[childIdx].triCount == 0) stack[stackPtr++] = childIdx;
		}
		if (stackPtr == 0) break;
		nodeIdx = stack[--stackPtr];
	}
	if (m[0].triCount > 0) /* root is a leaf: extra level :5036 */
	{
		m[1] = m[0];
		m[0].childCount = 1, m[0].child[0] = 1, m[0].triCount = 0;
	}
}

/* BVH8_CWBVH::ConvertFrom :5884.  data: room for triCount * 5 float4 blocks (zeroed here), tris: idxCount * 3 float4 (zeroed here). */
static uint32_t cwbvh_encode( mnode* m, const uint32_t* primIdx, const float* verts, float* data, float* tris, const uint32_t triCount, const uint32_t idxCount )
{
	memset( data, 0, (size_t)triCount * 5 * 16 ), memset( tris, 0, (size_t)idxCount * 3 * 16 );
	mnode* stackNodePtr[256];
	uint32_t stackNodeAddr[256], stackPtr = 1, nodeDataPtr = 5, triDataPtr = 0;
	stackNodePtr[0] = &m[0], stackNodeAddr[0] = 0;
	while (stackPtr > 0)
	{
		mnode* orig = stackNodePtr[--stackPtr];
		const int32_t currentNodeAddr = (int32_t)stackNodeAddr[stackPtr];
		const float* nodeLo = orig->mn, * nodeHi = orig->mx;
		const float nodeCentroid[3] = { (nodeLo[0] + nodeHi[0]) * 0.5f, (nodeLo[1] + nodeHi[1]) * 0.5f, (nodeLo[2] + nodeHi[2]) * 0.5f };
		float cost[8][8];
		int32_t assignment[8], isSlotEmpty[8];
		for (int32_t s = 0; s < 8; s++)
		{
			isSlotEmpty[s] = 1, assignment[s] = -1;
			const float ds[3] = { ((s >> 2) & 1) ? -1.0f : 1.0f, ((s >> 1) & 1) ? -1.0f : 1.0f, (s & 1) ? -1.0f : 1.0f };
			for (int32_t i = 0; i < 8; i++) if (orig->child[i] == 0) cost[s][i] = BVH_FAR; else
			{
				const mnode* child = &m[orig->child[i]];
				const float d[3] = { (child->mn[0] + child->mx[0]) * 0.5f - nodeCentroid[0], (child->mn[1] + child->mx[1]) * 0.5f - nodeCentroid[1], (child->mn[2] + child->mx[2]) * 0.5f - nodeCentroid[2] };
				cost[s][i] = fmaf( d[2], ds[2], fmaf( d[0], ds[0], d[1] * ds[1] ) ); /* tinybvh_dot; ds = +-1, so only the order of the additions matters */
			}
		}
		while (1)
		{
			float minCost = BVH_FAR;
			int32_t minEntryx = -1, minEntryy = -1;
			for (int32_t s = 0; s < 8; s++) for (int32_t i = 0; i < 8; i++)
				if (assignment[i] == -1 && isSlotEmpty[s] && cost[s][i] < minCost) minCost = cost[s][i], minEntryx = s, minEntryy = i;
			if (minEntryx == -1 && minEntryy == -1) break;
			isSlotEmpty[minEntryx] = 0, assignment[minEntryy] = minEntryx;
		}
		for (int32_t i = 0; i < 8; i++) if (assignment[i] == -1) for (int32_t s = 0; s < 8; s++) if (isSlotEmpty[s]) { isSlotEmpty[s] = 0, assignment[i] = s; break; }
		const mnode oldNode = *orig;
		for (int32_t i = 0; i < 8; i++) orig->child[assignment[i]] = oldNode.child[i];
		const int32_t ex = (int32_t)((int8_t)ceilf( log2f( (nodeHi[0] - nodeLo[0]) / 255.0f ) ));
		const int32_t ey = (int32_t)((int8_t)ceilf( log2f( (nodeHi[1] - nodeLo[1]) / 255.0f ) ));
		const int32_t ez = (int32_t)((int8_t)ceilf( log2f( (nodeHi[2] - nodeLo[2]) / 255.0f ) ));
		int32_t internalChildCount = 0, leafChildTriCount = 0, childBaseIndex = 0, triangleBaseIndex = 0;
		uint8_t imask = 0;
		for (int32_t i = 0; i < 8; i++)
		{
			if (orig->child[i] == 0) continue;
			mnode* child = &m[orig->child[i]];
			const int32_t qlox = (int32_t)floorf( (child->mn[0] - nodeLo[0]) / powf( 2, (float)ex ) );
			const int32_t qloy = (int32_t)floorf( (child->mn[1] - nodeLo[1]) / powf( 2, (float)ey ) );
			const int32_t qloz = (int32_t)floorf( (child->mn[2] - nodeLo[2]) / powf( 2, (float)ez ) );
			const int32_t qhix = (int32_t)ceilf( (child->mx[0] - nodeLo[0]) / powf( 2, (float)ex ) );
			const int32_t qhiy = (int32_t)ceilf( (child->mx[1] - nodeLo[1]) / powf( 2, (float)ey ) );
			const int32_t qhiz = (int32_t)ceilf( (child->mx[2] - nodeLo[2]) / powf( 2, (float)ez ) );
			uint8_t* baseAddr = (uint8_t*)&data[(size_t)(currentNodeAddr + 2) * 4];
			baseAddr[i + 0] = (uint8_t)qlox, baseAddr[i + 24] = (uint8_t)qhix;
			baseAddr[i + 8] = (uint8_t)qloy, baseAddr[i + 32] = (uint8_t)qhiy;
			baseAddr[i + 16] = (uint8_t)qloz, baseAddr[i + 40] = (uint8_t)qhiz;
			uint8_t* childMetaField = ((uint8_t*)&data[(size_t)(currentNodeAddr + 1) * 4]) + 8;
			if (child->triCount == 0)
			{
				const int32_t childNodeAddr = (int32_t)nodeDataPtr;
				if (internalChildCount++ == 0) childBaseIndex = childNodeAddr / 5;
				nodeDataPtr += 5, imask |= (uint8_t)(1 << i);
				childMetaField[i] = (uint8_t)((1 << 5) | (24 + (uint8_t)i));
				stackNodePtr[stackPtr] = child, stackNodeAddr[stackPtr++] = (uint32_t)childNodeAddr;
				internalChildCount++;
				continue;
			}
			const uint32_t tcount = child->triCount;
			if (leafChildTriCount == 0) triangleBaseIndex = (int32_t)triDataPtr;
			const int32_t unary = tcount == 1 ? 1 : tcount == 2 ? 3 : 7;
			childMetaField[i] = (uint8_t)((unary << 5) | leafChildTriCount);
			leafChildTriCount += (int32_t)tcount;
			for (uint32_t j = 0; j < tcount; j++)
			{
				const uint32_t triIdx = primIdx[child->firstTri + j];
				const float* v0 = verts + (size_t)triIdx * 12, * v1 = v0 + 4, * v2 = v0 + 8;
				float* o = tris + (size_t)triDataPtr * 4;
				for (int k = 0; k < 4; k++) o[k] = v2[k] - v0[k], o[4 + k] = v1[k] - v0[k]; /* bvhvec4 subtraction: the w lane too */
				o[8] = v0[0], o[9] = v0[1], o[10] = v0[2];
				memcpy( o + 11, &triIdx, 4 );
				triDataPtr += 3;
			}
		}
		float* n0 = data + (size_t)currentNodeAddr * 4;
		n0[0] = nodeLo[0], n0[1] = nodeLo[1], n0[2] = nodeLo[2];
		const uint8_t exyz[4] = { (uint8_t)ex, (uint8_t)ey, (uint8_t)ez, imask };
		memcpy( n0 + 3, exyz, 4 );
		memcpy( data + (size_t)(currentNodeAddr + 1) * 4, &childBaseIndex, 4 );
		memcpy( data + (size_t)(currentNodeAddr + 1) * 4 + 1, &triangleBaseIndex, 4 );
	}
	return nodeDataPtr;
}

/* The conversion chain of BVH8_CWBVH::Build / BuildHQ applied to a BVH2: SplitLeafs( 3 ), MBVH<8>::ConvertFrom, BVH8_CWBVH::ConvertFrom.
 * nodes / usedNodes / primIdx / idxCount: the source tree (not modified); triCount: BVHBase::triCount of the source.
 * data: triCount * 5 float4, tris: idxCount * 3 float4.  Returns usedBlocks. */
uint32_t orc_cwbvh_from_bvh( const orc_node* nodes, uint32_t usedNodes, const uint32_t* primIdx, uint32_t idxCount, const float* verts, uint32_t triCount, float* data, float* tris )
{
	/* SplitLeafs can add two nodes per excess leaf chunk: bounded by the number of index entries */
	const size_t room = (size_t)usedNodes + 2 * (size_t)idxCount + 4;
	orc_node* b = (orc_node*)calloc( room, sizeof( orc_node ) );
	memcpy( b, nodes, (size_t)usedNodes * sizeof( orc_node ) );
	const uint32_t used = split_leafs( b, usedNodes, 3 );
	mnode* m = (mnode*)calloc( room, sizeof( mnode ) );
	mbvh8_from_bvh( b, used, m );
	const uint32_t blocks = cwbvh_encode( m, primIdx, verts, data, tris, triCount, idxCount );
	free( b ), free( m );
	return blocks;
}

/* ---- BVH8_CWBVH::Intersect :7046-7154 */
static inline uint32_t bfind( uint32_t v ) { return 31u - (uint32_t)__builtin_clz( v ); }
static inline uint32_t sign_extend_s8x4( uint32_t i ) { return ((i & 0x80000000u) ? 0xff000000u : 0) + ((i & 0x00800000u) ? 0x00ff0000u : 0) + ((i & 0x00008000u) ? 0x0000ff00u : 0) + ((i & 0x00000080u) ? 0x000000ffu : 0); }
static inline float fmax_( float a, float b ) { return a > b ? a : b; }
static inline float fmin_( float a, float b ) { return a < b ? a : b; }
typedef struct { float O[3]; uint32_t mask; float D[3]; uint32_t instIdx; float rD[3]; uint32_t pad; float t, u, v; uint32_t prim; uint8_t aux[64]; } cw_ray;

static uint32_t slab4( uint32_t meta4, uint32_t octinv, uint32_t lox, uint32_t loy, uint32_t loz, uint32_t hix, uint32_t hiy, uint32_t hiz,
	const float* adj, const float* org, float tmin, float tmax )
{
	const uint32_t is_inner4 = (meta4 & (meta4 << 1)) & 0x10101010u;
	const uint32_t inner_mask4 = sign_extend_s8x4( is_inner4 << 3 );
	const uint32_t bit_index4 = (meta4 ^ (octinv & inner_mask4)) & 0x1F1F1F1Fu;
	const uint32_t child_bits4 = (meta4 >> 5) & 0x07070707u;
	uint32_t hitmask = 0;
	for (int i = 0; i < 4; i++)
	{
		const float tminx = fmaf( (float)((lox >> (8 * i)) & 0xFF), adj[0], org[0] ), tminy = fmaf( (float)((loy >> (8 * i)) & 0xFF), adj[1], org[1] ), tminz = fmaf( (float)((loz >> (8 * i)) & 0xFF), adj[2], org[2] );
		const float tmaxx = fmaf( (float)((hix >> (8 * i)) & 0xFF), adj[0], org[0] ), tmaxy = fmaf( (float)((hiy >> (8 * i)) & 0xFF), adj[1], org[1] ), tmaxz = fmaf( (float)((hiz >> (8 * i)) & 0xFF), adj[2], org[2] );
		const float cmin = fmax_( fmax_( fmax_( tminx, tminy ), tminz ), tmin );
		const float cmax = fmin_( fmin_( fmin_( tmaxx, tmaxy ), tmaxz ), tmax );
		if (cmin <= cmax) hitmask |= ((child_bits4 >> (8 * i)) & 0xFF) << ((bit_index4 >> (8 * i)) & 0xFF);
	}
	return hitmask;
}

static void cw_intersect1( const float* nodes, const float* tris, cw_ray* ray )
{
	uint32_t stack[128][2], stackPtr = 0, hitAddr = 0;
	float tu = 0, tv = 0;
	const float tmin = 0;
	float tmax = ray->t;
	const uint32_t octinv = (7u - ((ray->D[0] < 0 ? 4u : 0u) | (ray->D[1] < 0 ? 2u : 0u) | (ray->D[2] < 0 ? 1u : 0u))) * 0x1010101u;
	uint32_t ngx = 0, ngy = 0x80000000u, tgx = 0, tgy = 0;
	while (1)
	{
		if (ngy > 0x00FFFFFFu)
		{
			const uint32_t hits = ngy, imask = ngy;
			const uint32_t child_bit_index = bfind( hits ), child_node_base_index = ngx;
			ngy &= ~(1u << child_bit_index);
			if (ngy > 0x00FFFFFFu) stack[stackPtr][0] = ngx, stack[stackPtr++][1] = ngy;
			const uint32_t slot_index = (child_bit_index - 24) ^ (octinv & 255);
			const uint32_t relative_index = (uint32_t)__builtin_popcount( imask & ~(0xFFFFFFFFu << slot_index) );
			const float* n = nodes + (size_t)(child_node_base_index + relative_index) * 20;
			const uint32_t* w = (const uint32_t*)n;
			const int8_t* e = (const int8_t*)&n[3];
			ngx = w[4], tgx = w[5], tgy = 0;
			const uint32_t vx = (uint32_t)(e[0] + 127) << 23, vy = (uint32_t)(e[1] + 127) << 23, vz = (uint32_t)(e[2] + 127) << 23;
			float sx, sy, sz;
			memcpy( &sx, &vx, 4 ), memcpy( &sy, &vy, 4 ), memcpy( &sz, &vz, 4 );
			const float adj[3] = { sx * ray->rD[0], sy * ray->rD[1], sz * ray->rD[2] };
			const float org[3] = { -(ray->O[0] - n[0]) * ray->rD[0], -(ray->O[1] - n[1]) * ray->rD[1], -(ray->O[2] - n[2]) * ray->rD[2] };
			const int nx = ray->rD[0] < 0, ny = ray->rD[1] < 0, nz = ray->rD[2] < 0;
			/* words: 8,9 = qlox; 10,11 = qloy; 12,13 = qloz; 14,15 = qhix; 16,17 = qhiy; 18,19 = qhiz */
			uint32_t hitmask = slab4( w[6], octinv, nx ? w[14] : w[8], ny ? w[16] : w[10], nz ? w[18] : w[12], nx ? w[8] : w[14], ny ? w[10] : w[16], nz ? w[12] : w[18], adj, org, tmin, tmax );
			hitmask |= slab4( w[7], octinv, nx ? w[15] : w[9], ny ? w[17] : w[11], nz ? w[19] : w[13], nx ? w[9] : w[15], ny ? w[11] : w[17], nz ? w[13] : w[19], adj, org, tmin, tmax );
			ngy = (hitmask & 0xFF000000u) | (w[3] >> 24), tgy = hitmask & 0x00FFFFFFu;
		}
		else tgx = ngx, tgy = ngy, ngx = 0, ngy = 0;
		while (tgy != 0)
		{
			const uint32_t triangleIndex = bfind( tgy );
			tgy -= 1u << triangleIndex;
			const float* tp = tris + (size_t)(tgx + triangleIndex * 3) * 4; /* e2, e1, v0 | prim */
			const float* e2 = tp, * e1 = tp + 4, * v0 = tp + 8;
			/* MOLLER_TRUMBORE_TEST :1644 in the pairing of tbvh_oracle.c */
			const float* D = ray->D, * O = ray->O;
			const float hx = fmaf( D[1], e2[2], -(D[2] * e2[1]) ), hy = fmaf( D[2], e2[0], -(D[0] * e2[2]) ), hz = fmaf( D[0], e2[1], -(D[1] * e2[0]) );
			const float a = fmaf( e1[2], hz, fmaf( e1[0], hx, e1[1] * hy ) );
			if (fabsf( a ) < 0.000001f) continue;
			const float f = 1 / a;
			const float sx = O[0] - v0[0], sy = O[1] - v0[1], sz = O[2] - v0[2];
			const float u = f * fmaf( hz, sz, fmaf( hx, sx, hy * sy ) );
			const float qx = fmaf( -e1[1], sz, e1[2] * sy ), qy = fmaf( -e1[2], sx, e1[0] * sz ), qz = fmaf( -e1[0], sy, e1[1] * sx );
			const float v = f * fmaf( D[2], qz, fmaf( D[1], qy, D[0] * qx ) );
			if (u < 0 || v < 0 || u + v > 1) continue;
			const float t = f * fmaf( e2[2], qz, fmaf( e2[0], qx, e2[1] * qy ) );
			if (t < 0 || t > tmax) continue;
			tu = u, tv = v, tmax = t;
			memcpy( &hitAddr, tp + 11, 4 );
		}
		if (ngy > 0x00FFFFFFu) continue;
		if (stackPtr > 0) { stackPtr--; ngx = stack[stackPtr][0], ngy = stack[stackPtr][1]; }
		else
		{
			ray->t = tmax;
			if (tmax < BVH_FAR) ray->u = tu, ray->v = tv, ray->prim = hitAddr;
			break;
		}
	}
}

void orc_cwbvh_intersect( const float* bvh8Data, const float* bvh8Tris, void* rays, uint64_t n )
{
	for (uint64_t i = 0; i < n; i++) cw_intersect1( bvh8Data, bvh8Tris, (cw_ray*)rays + i );
}

/* ---- TLAS over CWBVH BLASses: see tbvh_oracle.h */
static int walk_cw_blas( const void* user, uint32_t blasIdx, void* temp_, int anyhit )
{
	const orc_cwblas* b = (const orc_cwblas*)user + blasIdx;
	cw_ray* temp = (cw_ray*)temp_;
	const cw_ray before = *temp;
	cw_intersect1( b->bvh8Data, b->bvh8Tris, temp );
	if (temp->t < before.t)
	{
		temp->pad = temp->instIdx; /* hit.inst: the instance travels with the hit (traverse_tlas.cl:88; IntersectTri :8525) */
		return anyhit;
	}
	temp->t = before.t, temp->u = before.u, temp->v = before.v, temp->prim = before.prim, temp->pad = before.pad;
	return 0;
}
void orc_intersect_tlas_cw( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_cwblas* blas, void* rays, uint64_t n )
{
	for (uint64_t i = 0; i < n; i++) orc_tlas_walk1( nodes, primIdx, inst, (cw_ray*)rays + i, 0, walk_cw_blas, blas );
}
void orc_occluded_tlas_cw( const orc_node* nodes, const uint32_t* primIdx, const orc_instance* inst, const orc_cwblas* blas, const void* rays, uint64_t n, uint32_t* bits )
{
	for (uint64_t i = 0; i < n; i++)
	{
		cw_ray tmp = ((const cw_ray*)rays)[i];
		if (orc_tlas_walk1( nodes, primIdx, inst, &tmp, 1, walk_cw_blas, blas )) bits[i >> 5] |= 1u << (i & 31);
	}
}
