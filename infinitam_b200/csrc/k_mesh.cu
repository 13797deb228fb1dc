// Marching cubes over the voxel-block hash.
//
// Replaces ITMMeshingEngine::MeshScene
//   CPU driver     ITMLib/Engine/DeviceSpecific/CPU/ITMMeshingEngine_CPU.cpp:19-58
//   per voxel      findPointNeighbors / sdfInterp / buildVertList  ITMLib/Engine/DeviceAgnostic/ITMMeshingEngine.h:153-232
//
// The reference walks the hash table serially (entry id ascending, then z, y, x, then the cube's triangles) and appends
// to one array, so triangle i's position in the output is the number of triangles every earlier voxel produced.  The same
// array - bit for bit, in the same order - is produced here in four launches:
//   1. k_find_visible's ordered scan with the predicate "allocated" compacts the entries that own a voxel block
//      (ascending entry id; at most SDF_LOCAL_BLOCK_NUM of them),
//   2. k_mesh_blocks<COUNT>: one 512-thread CTA per listed block, thread = voxel in (z, y, x) order; the block's 9x9x9
//      SDF corner lattice (own block + the 7 neighbours towards +x/+y/+z, found by hash lookup) is staged in shared
//      memory once, every thread classifies its cube and the CTA publishes its triangle count,
//   3. k_mesh_scan: exclusive prefix over the per-block counts (one CTA, sequential over chunks: at most 64 values per
//      thread at SDF_LOCAL_BLOCK_NUM = 65536),
//   4. k_mesh_blocks<EMIT>: same staging, CTA-wide exclusive scan of the per-voxel counts, triangles written at
//      base + offset.
// The reference's overflow rule (the write index stops at noMaxTriangles - 1, so that slot ends up holding the last
// triangle emitted) is kept: triangles beyond the capacity are dropped except the very last one.
//
// The case table is the classic Lorensen-Cline / Bourke marching-cubes triangulation (the reference uses the same one,
// ITMMeshingEngine.h:9-151), stored here as one 64-bit word per case: nibble i = i-th edge index, 0xF terminates.  The
// 12-bit edge mask of a case is the union of the edges its triangles use, so no second table is needed.
// tests/test_mc_tables.py checks both against the reference's tables when the reference tree is present.
#include "itm_common.cuh"
#include "kernels.h"

namespace {

using namespace itm;

__constant__ unsigned long long MC_CASE[256] = {
    0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFF380ULL, 0xFFFFFFFFFFFFF910ULL, 0xFFFFFFFFFF189381ULL,
    0xFFFFFFFFFFFFFA21ULL, 0xFFFFFFFFFFA21380ULL, 0xFFFFFFFFFF920A29ULL, 0xFFFFFFF89A8A2382ULL,
    0xFFFFFFFFFFFFF2B3ULL, 0xFFFFFFFFFF0B82B0ULL, 0xFFFFFFFFFFB32091ULL, 0xFFFFFFFB89B912B1ULL,
    0xFFFFFFFFFF3AB1A3ULL, 0xFFFFFFFAB8A801A0ULL, 0xFFFFFFF9AB9B3093ULL, 0xFFFFFFFFFFB8AA89ULL,
    0xFFFFFFFFFFFFF874ULL, 0xFFFFFFFFFF437034ULL, 0xFFFFFFFFFF748910ULL, 0xFFFFFFF137174914ULL,
    0xFFFFFFFFFF748A21ULL, 0xFFFFFFFA21403743ULL, 0xFFFFFFF748209A29ULL, 0xFFFF4973727929A2ULL,
    0xFFFFFFFFFF2B3748ULL, 0xFFFFFFF40242B74BULL, 0xFFFFFFFB32748109ULL, 0xFFFF1292B9B49B74ULL,
    0xFFFFFFF487AB31A3ULL, 0xFFFF4B7401B41AB1ULL, 0xFFFF30BAB9B09874ULL, 0xFFFFFFFAB99B4B74ULL,
    0xFFFFFFFFFFFFF459ULL, 0xFFFFFFFFFF380459ULL, 0xFFFFFFFFFF051450ULL, 0xFFFFFFF513538458ULL,
    0xFFFFFFFFFF459A21ULL, 0xFFFFFFF594A21803ULL, 0xFFFFFFF204245A25ULL, 0xFFFF8434535235A2ULL,
    0xFFFFFFFFFFB32459ULL, 0xFFFFFFF594B802B0ULL, 0xFFFFFFFB32510450ULL, 0xFFFF584B82852512ULL,
    0xFFFFFFF45931AB3AULL, 0xFFFFAB81A8180594ULL, 0xFFFF30BAB5B05045ULL, 0xFFFFFFFB8AA85845ULL,
    0xFFFFFFFFFF975879ULL, 0xFFFFFFF375359039ULL, 0xFFFFFFF751710870ULL, 0xFFFFFFFFFF753351ULL,
    0xFFFFFFF21A759879ULL, 0xFFFF37503505921AULL, 0xFFFF25A758528208ULL, 0xFFFFFFF7533525A2ULL,
    0xFFFFFFF2B3987597ULL, 0xFFFFB72029279759ULL, 0xFFFF751871810B32ULL, 0xFFFFFFF51771B12BULL,
    0xFFFFB3A31A758859ULL, 0xF0ABA010B7905075ULL, 0xF07570805A30B0ABULL, 0xFFFFFFFFFF5B75ABULL,
    0xFFFFFFFFFFFFF56AULL, 0xFFFFFFFFFF6A5380ULL, 0xFFFFFFFFFF6A5109ULL, 0xFFFFFFF6A5891381ULL,
    0xFFFFFFFFFF162561ULL, 0xFFFFFFF803621561ULL, 0xFFFFFFF620609569ULL, 0xFFFF823625285895ULL,
    0xFFFFFFFFFF56AB32ULL, 0xFFFFFFF56A02B80BULL, 0xFFFFFFF6A5B32910ULL, 0xFFFFB892B92916A5ULL,
    0xFFFFFFF315356B36ULL, 0xFFFF6B51505B0B80ULL, 0xFFFF9505606306B3ULL, 0xFFFFFFF89BB96956ULL,
    0xFFFFFFFFFF8746A5ULL, 0xFFFFFFFA56374034ULL, 0xFFFFFFF7486A5091ULL, 0xFFFF49737179156AULL,
    0xFFFFFFF874156216ULL, 0xFFFF743403625521ULL, 0xFFFF620560509748ULL, 0xF962695923497937ULL,
    0xFFFFFFF56A4872B3ULL, 0xFFFFB720242746A5ULL, 0xFFFF6A5B32874910ULL, 0xF6A54B7B492B9129ULL,
    0xFFFF6B51535B3748ULL, 0xFB404B7B016B5B15ULL, 0xF74836B630560950ULL, 0xFFFF9B7974B96956ULL,
    0xFFFFFFFFFFA4694AULL, 0xFFFFFFF380A946A4ULL, 0xFFFFFFF04606A10AULL, 0xFFFFA16468618138ULL,
    0xFFFFFFF462421941ULL, 0xFFFF462942921803ULL, 0xFFFFFFFFFF624420ULL, 0xFFFFFFF624428238ULL,
    0xFFFFFFF32B46A94AULL, 0xFFFF6A4A94B82280ULL, 0xFFFFA164606102B3ULL, 0xF1B8B12184A16146ULL,
    0xFFFF36B319639469ULL, 0xF14641916B0181B8ULL, 0xFFFFFFF4600636B3ULL, 0xFFFFFFFFFF86B846ULL,
    0xFFFFFFFA98A876A7ULL, 0xFFFFA76A907A0370ULL, 0xFFFF0818717A176AULL, 0xFFFFFFF37117A76AULL,
    0xFFFF768981861621ULL, 0xF937390976192962ULL, 0xFFFFFFF206607087ULL, 0xFFFFFFFFFF276237ULL,
    0xFFFF76898A86AB32ULL, 0xF7A9A76790B72702ULL, 0xFB32A767A1871081ULL, 0xFFFF17616A71B12BULL,
    0xF63136B619768698ULL, 0xFFFFFFFFFF76B190ULL, 0xFFFF06B0B3607087ULL, 0xFFFFFFFFFFFFF6B7ULL,
    0xFFFFFFFFFFFFFB67ULL, 0xFFFFFFFFFF67B803ULL, 0xFFFFFFFFFF67B910ULL, 0xFFFFFFF67B138918ULL,
    0xFFFFFFFFFF7B621AULL, 0xFFFFFFF7B6803A21ULL, 0xFFFFFFF7B69A2092ULL, 0xFFFF89A38A3A27B6ULL,
    0xFFFFFFFFFF726327ULL, 0xFFFFFFF026067807ULL, 0xFFFFFFF910732672ULL, 0xFFFF678891681261ULL,
    0xFFFFFFF73171A67AULL, 0xFFFF801781A7167AULL, 0xFFFF7A69A0A70730ULL, 0xFFFFFFF9A88A7A67ULL,
    0xFFFFFFFFFF68B486ULL, 0xFFFFFFF640603B63ULL, 0xFFFFFFF109648B68ULL, 0xFFFF63B139369649ULL,
    0xFFFFFFF1A28B6486ULL, 0xFFFF640B60B03A21ULL, 0xFFFF9A2920B648B4ULL, 0xF36463B34923A39AULL,
    0xFFFFFFF264248328ULL, 0xFFFFFFFFFF264240ULL, 0xFFFF834642432091ULL, 0xFFFFFFF642241491ULL,
    0xFFFF1A6648168318ULL, 0xFFFFFFF40660A01AULL, 0xF39A9303A6834364ULL, 0xFFFFFFFFFF4A649AULL,
    0xFFFFFFFFFFB67594ULL, 0xFFFFFFF67B594380ULL, 0xFFFFFFFB67045105ULL, 0xFFFF51345343867BULL,
    0xFFFFFFFB6721A459ULL, 0xFFFF594380A217B6ULL, 0xFFFF204A24A45B67ULL, 0xF67B25A523453843ULL,
    0xFFFFFFF945267327ULL, 0xFFFF786260680459ULL, 0xFFFF045051673263ULL, 0xF851584812786826ULL,
    0xFFFF73167161A459ULL, 0xF459078701671A61ULL, 0xFA737A6A305A4A04ULL, 0xFFFFA84A458A7A67ULL,
    0xFFFFFFF98B9B6596ULL, 0xFFFF590650360B63ULL, 0xFFFFB65510B508B0ULL, 0xFFFFFFF1355363B6ULL,
    0xFFFF65B8B9B59A21ULL, 0xFA21965690B603B0ULL, 0xF52025A50865B58BULL, 0xFFFF35A3A25363B6ULL,
    0xFFFF283265825985ULL, 0xFFFFFFF260069659ULL, 0xF826283865081851ULL, 0xFFFFFFFFFF612651ULL,
    0xF698965683A61631ULL, 0xFFFF06505960A01AULL, 0xFFFFFFFFFFA65830ULL, 0xFFFFFFFFFFFFF65AULL,
    0xFFFFFFFFFFB57A5BULL, 0xFFFFFFF03857BA5BULL, 0xFFFFFFF091BA57B5ULL, 0xFFFF1381897BA57AULL,
    0xFFFFFFF15717B21BULL, 0xFFFFB27571721380ULL, 0xFFFF7B2209729579ULL, 0xF289823295B27257ULL,
    0xFFFFFFF573532A52ULL, 0xFFFF52A578258028ULL, 0xFFFF2A37353A5109ULL, 0xF25752A278129289ULL,
    0xFFFFFFFFFF573531ULL, 0xFFFFFFF571170780ULL, 0xFFFFFFF735539309ULL, 0xFFFFFFFFFF795789ULL,
    0xFFFFFFF8BA8A5485ULL, 0xFFFF03BBA50B5405ULL, 0xFFFF54ABA8A48910ULL, 0xF41314943B54A4BAULL,
    0xFFFF8548B2582152ULL, 0xFB151B2B543B0B40ULL, 0xF58B8545B2950520ULL, 0xFFFFFFFFFF3B2549ULL,
    0xFFFF483543253A52ULL, 0xFFFFFFF0244252A5ULL, 0xF910854583A532A3ULL, 0xFFFF2492914252A5ULL,
    0xFFFFFFF153358548ULL, 0xFFFFFFFFFF501540ULL, 0xFFFF530509358548ULL, 0xFFFFFFFFFFFFF549ULL,
    0xFFFFFFFBA9B947B4ULL, 0xFFFFBA97B9794380ULL, 0xFFFFB470414B1BA1ULL, 0xF4BAB474A1843413ULL,
    0xFFFF219B294B97B4ULL, 0xF3801B2B197B9479ULL, 0xFFFFFFF04224B47BULL, 0xFFFF42343824B47BULL,
    0xFFFF947732972A92ULL, 0xF70207872A4797A9ULL, 0xFA040A1A472A3A73ULL, 0xFFFFFFFFFF4782A1ULL,
    0xFFFFFFF317714194ULL, 0xFFFF178180714194ULL, 0xFFFFFFFFFF347304ULL, 0xFFFFFFFFFFFFF784ULL,
    0xFFFFFFFFFF8BA8A9ULL, 0xFFFFFFFA9BB93903ULL, 0xFFFFFFFBA88A0A10ULL, 0xFFFFFFFFFFA3BA13ULL,
    0xFFFFFFF8B99B1B21ULL, 0xFFFF9B2921B93903ULL, 0xFFFFFFFFFFB08B20ULL, 0xFFFFFFFFFFFFFB23ULL,
    0xFFFFFFF98AA82832ULL, 0xFFFFFFFFFF2902A9ULL, 0xFFFF8A1810A82832ULL, 0xFFFFFFFFFFFFF2A1ULL,
    0xFFFFFFFFFF819831ULL, 0xFFFFFFFFFFFFF190ULL, 0xFFFFFFFFFFFFF830ULL, 0xFFFFFFFFFFFFFFFFULL,
};

// cube corner k -> (dx, dy, dz), in findPointNeighbors' order; edge e -> its two corners, in buildVertList's order
__constant__ unsigned char MC_EDGE_A[12] = {0, 1, 2, 3, 4, 5, 6, 7, 0, 1, 2, 3};
__constant__ unsigned char MC_EDGE_B[12] = {1, 2, 3, 0, 5, 6, 7, 4, 4, 5, 6, 7};

#define LAT 9  // corner lattice edge: 8 voxels + 1

struct MeshTri {
  float v[9];  // ITMMesh::Triangle {Vector3f p0, p1, p2}
};

__device__ __forceinline__ int find_block_ptr(const HashEntry *__restrict__ table, int bx, int by, int bz, int nBuckets, unsigned hashMask) {
  int hashIdx = (int)hash_index(bx, by, bz, hashMask);
  while (true) {
    const HashEntry e = load_entry(table, hashIdx);
    if (e.px == bx && e.py == by && e.pz == bz && e.ptr >= 0) return e.ptr;
    if (e.offset < 1) return -1;
    hashIdx = nBuckets + e.offset - 1;
  }
}

// sdfInterp (ITMMeshingEngine.h:194-201) on one coordinate triple
__device__ __forceinline__ void sdf_interp(const float *p1, const float *p2, float v1, float v2, float *out) {
  if (fabsf(0.0f - v1) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  if (fabsf(0.0f - v2) < 0.00001f) { out[0] = p2[0]; out[1] = p2[1]; out[2] = p2[2]; return; }
  if (fabsf(v1 - v2) < 0.00001f) { out[0] = p1[0]; out[1] = p1[1]; out[2] = p1[2]; return; }
  const float t = (0.0f - v1) / (v2 - v1);
  out[0] = p1[0] + t * (p2[0] - p1[0]);
  out[1] = p1[1] + t * (p2[1] - p1[1]);
  out[2] = p1[2] + t * (p2[2] - p1[2]);
}

template <int VW, bool EMIT>
__global__ void __launch_bounds__(512) k_mesh_blocks(const uint32_t *__restrict__ voxels, const HashEntry *__restrict__ table,
                                                     const int *__restrict__ blockList, const FrameState *__restrict__ st,
                                                     unsigned *__restrict__ counts, const unsigned long long *__restrict__ offsets,
                                                     MeshTri *__restrict__ triangles, unsigned noMaxTriangles, SceneParams sp) {
  __shared__ short sSdf[LAT * LAT * LAT];
  __shared__ int sPtr[8];
  __shared__ unsigned sWarp[16];
  const int nBlocks = st->noVisibleEntries;  // length of the compacted list
  const int tid = threadIdx.x;
  const int lx = tid & 7, ly = (tid >> 3) & 7, lz = tid >> 6;
  for (int b = blockIdx.x; b < nBlocks; b += gridDim.x) {
    const HashEntry e = load_entry(table, __ldg(blockList + b));
    if (tid < 8) {
      // neighbour n = (dx, dy, dz) bits; n == 0 is the block itself
      sPtr[tid] = tid == 0 ? e.ptr : find_block_ptr(table, e.px + (tid & 1), e.py + ((tid >> 1) & 1), e.pz + (tid >> 2), sp.nBuckets, sp.hashMask);
    }
    __syncthreads();
    // stage the 9^3 lattice: missing voxels read as "not found", which findPointNeighbors treats like sdf == 1.0f (raw 32767)
    for (int i = tid; i < LAT * LAT * LAT; i += 512) {
      const int x = i % LAT, y = (i / LAT) % LAT, z = i / (LAT * LAT);
      const int n = (x >> 3) | ((y >> 3) << 1) | ((z >> 3) << 2);
      const int ptr = sPtr[n];
      short v = 32767;
      if (ptr >= 0) v = (short)(__ldg(voxels + ((size_t)ptr * ITM_BLOCK_SIZE3 + (x & 7) + ((y & 7) << 3) + ((z & 7) << 6)) * VW) & 0xFFFFu);
      sSdf[i] = v;
    }
    __syncthreads();
    // classify this thread's cube; corner order of findPointNeighbors: (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)
    const int base = lx + ly * LAT + lz * LAT * LAT;
    short raw[8];
    raw[0] = sSdf[base];                 raw[1] = sSdf[base + 1];
    raw[2] = sSdf[base + 1 + LAT];       raw[3] = sSdf[base + LAT];
    raw[4] = sSdf[base + LAT * LAT];     raw[5] = sSdf[base + 1 + LAT * LAT];
    raw[6] = sSdf[base + 1 + LAT + LAT * LAT]; raw[7] = sSdf[base + LAT + LAT * LAT];
    bool valid = true;
    int cubeIndex = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      if (raw[k] == 32767) valid = false;
      if (raw[k] < 0) cubeIndex |= 1 << k;
    }
    unsigned long long tris = valid ? MC_CASE[cubeIndex] : ~0ull;
    unsigned nTri = 0;
    {
      unsigned long long t = tris;
      while ((t & 0xFull) != 0xFull) { nTri++; t >>= 12; }
    }
    // CTA-wide exclusive scan of nTri in thread order (= the serial loop's z, y, x order)
    const int lane = tid & 31, warp = tid >> 5;
    unsigned inc = nTri;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
      if (lane >= o) inc += v;
    }
    if (lane == 31) sWarp[warp] = inc;
    __syncthreads();
    unsigned warpBase = 0, total = 0;
#pragma unroll
    for (int w = 0; w < 16; ++w) {
      const unsigned s = sWarp[w];
      if (w < warp) warpBase += s;
      total += s;
    }
    if (!EMIT) {
      if (tid == 0) counts[b] = total;
    } else if (nTri) {
      const unsigned long long first = offsets[b] + warpBase + inc - nTri;
      const unsigned long long grand = offsets[nBlocks];
      // corner positions (voxel coordinates as floats) and SDF values as floats
      const int gx = e.px * ITM_BLOCK_SIZE + lx, gy = e.py * ITM_BLOCK_SIZE + ly, gz = e.pz * ITM_BLOCK_SIZE + lz;
      float vertList[12][3];
      unsigned edgeMask = 0;
      {
        unsigned long long t = tris;
        while ((t & 0xFull) != 0xFull) { edgeMask |= 1u << (unsigned)(t & 0xFull); t >>= 4; }
      }
#pragma unroll
      for (int ed = 0; ed < 12; ++ed) {
        if (!((edgeMask >> ed) & 1u)) continue;
        const int a = MC_EDGE_A[ed], c = MC_EDGE_B[ed];
        const int ax = ((a & 3) == 1 || (a & 3) == 2) ? 1 : 0, ay = (a & 2) ? 1 : 0, az = a >> 2;
        const int cx = ((c & 3) == 1 || (c & 3) == 2) ? 1 : 0, cy = (c & 2) ? 1 : 0, cz = c >> 2;
        const float pa[3] = {(float)(gx + ax), (float)(gy + ay), (float)(gz + az)};
        const float pc[3] = {(float)(gx + cx), (float)(gy + cy), (float)(gz + cz)};
        sdf_interp(pa, pc, (float)raw[a] / 32767.0f, (float)raw[c] / 32767.0f, vertList[ed]);
      }
      unsigned long long t = tris;
      for (unsigned i = 0; i < nTri; ++i, t >>= 12) {
        unsigned long long slot = first + i;
        if (slot >= (unsigned long long)noMaxTriangles - 1ull) {
          // beyond the capacity: the reference keeps overwriting slot noMax-1; only the last triangle of all survives there
          if (slot != grand - 1ull) continue;
          slot = (unsigned long long)noMaxTriangles - 1ull;
        }
        MeshTri out;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int ed = (int)((t >> (4 * c)) & 0xFull);
          out.v[c * 3 + 0] = vertList[ed][0] * sp.voxelSize;
          out.v[c * 3 + 1] = vertList[ed][1] * sp.voxelSize;
          out.v[c * 3 + 2] = vertList[ed][2] * sp.voxelSize;
        }
        triangles[slot] = out;
      }
    }
    __syncthreads();  // sSdf / sPtr / sWarp are reused by the next block
  }
}

// exclusive prefix of counts[0..n) into offsets[0..n], offsets[n] = total; one CTA
__global__ void __launch_bounds__(1024) k_mesh_scan(const unsigned *__restrict__ counts, unsigned long long *__restrict__ offsets,
                                                    FrameState *st, unsigned noMaxTriangles) {
  __shared__ unsigned long long sPart[1024];
  const int n = st->noVisibleEntries;
  const int per = (n + 1023) / 1024;
  const int lo = min(n, (int)threadIdx.x * per), hi = min(n, lo + per);
  unsigned long long sum = 0;
  for (int i = lo; i < hi; ++i) sum += counts[i];
  sPart[threadIdx.x] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long acc = 0;
    for (int t = 0; t < 1024; ++t) {
      const unsigned long long v = sPart[t];
      sPart[t] = acc;
      acc += v;
    }
    offsets[n] = acc;
    // mesh->noTotalTriangles: the write index never passes noMaxTriangles - 1
    st->noMeshTriangles = (int)(acc < (unsigned long long)noMaxTriangles - 1ull ? acc : (unsigned long long)noMaxTriangles - 1ull);
  }
  __syncthreads();
  unsigned long long run = sPart[threadIdx.x];
  for (int i = lo; i < hi; ++i) {
    offsets[i] = run;
    run += counts[i];
  }
}

}  // namespace

namespace itm {

void launch_mesh_scene(const MeshArgs &a, cudaStream_t s) {
  // mesh->triangles->Clear()
  cudaMemsetAsync(a.triangles, 0, (size_t)a.noMaxTriangles * sizeof(MeshTri), s);
  launch_find_visible_blocks(a.hashTable, a.blockList, a.st, ViewParams{0, 0, 0.f, 0.f, 0.f, 0.f}, a.sp, a.sp.nLocal, a.ticket, a.tileState, s,
                             /*allAllocated=*/1);
  const uint32_t *vox = reinterpret_cast<const uint32_t *>(a.voxels);
  const HashEntry *table = reinterpret_cast<const HashEntry *>(a.hashTable);
  MeshTri *tri = reinterpret_cast<MeshTri *>(a.triangles);
  const int grid = 148 * 4;
  if (a.sp.voxelWords == 2) k_mesh_blocks<2, false><<<grid, 512, 0, s>>>(vox, table, a.blockList, a.st, a.counts, a.offsets, tri, a.noMaxTriangles, a.sp);
  else k_mesh_blocks<1, false><<<grid, 512, 0, s>>>(vox, table, a.blockList, a.st, a.counts, a.offsets, tri, a.noMaxTriangles, a.sp);
  k_mesh_scan<<<1, 1024, 0, s>>>(a.counts, a.offsets, a.st, a.noMaxTriangles);
  if (a.sp.voxelWords == 2) k_mesh_blocks<2, true><<<grid, 512, 0, s>>>(vox, table, a.blockList, a.st, a.counts, a.offsets, tri, a.noMaxTriangles, a.sp);
  else k_mesh_blocks<1, true><<<grid, 512, 0, s>>>(vox, table, a.blockList, a.st, a.counts, a.offsets, tri, a.noMaxTriangles, a.sp);
}

}  // namespace itm
