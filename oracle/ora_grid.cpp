// ORACLE — test infrastructure only (see ora_math.h).  Photon uniform-grid build restated from
// RenderEngine/renderer/OptixRenderer_SpatialHash.cu:42-282 (thrust::sort_by_key is a stable radix sort,
// restated as std::stable_sort by key).
#include "ora_core.h"
#include <numeric>

namespace ora {

// getSmallestPossibleCellSize, OptixRenderer_SpatialHash.cu:62-71
static float getSmallestPossibleCellSize(const float3& sceneExtent, const unsigned maxGridSize)
{
    float sceneVolume = sceneExtent.x * sceneExtent.y * sceneExtent.z;
    float minVolumePerCell = sceneVolume / maxGridSize;
    float smallestPossibleRadiusC = powf(minVolumePerCell, 1.0f / 3.0f);
    float3 numCellsF = sceneExtent / smallestPossibleRadiusC;          // float3/float = * (1/s)
    uint3 numCells = {f2u_sat(floorf(numCellsF.x)), f2u_sat(floorf(numCellsF.y)), f2u_sat(floorf(numCellsF.z))};
    float3 radiusEachAxis = {sceneExtent.x / numCells.x, sceneExtent.y / numCells.y, sceneExtent.z / numCells.z};
    return fmaxf3(radiusEachAxis);
}

// calculateGridSize, :42-50
static uint3 calculateGridSize(const float3& sceneSize, const float radius)
{
    float3 f = sceneSize / radius;
    uint3 r = {f2u_sat(ceilf(f.x)), f2u_sat(ceilf(f.y)), f2u_sat(ceilf(f.z))};
    r.x = std::max(1u, r.x); r.y = std::max(1u, r.y); r.z = std::max(1u, r.z);
    return r;
}

bool build_photon_grid(Photon* photons, uint n, std::vector<uint>& keys, std::vector<uint>& offsets, PhotonGrid& g)
{
    // getPhotonsBoundingBox :123-128 with PhotonToAABBConverter :112-119 / AABBReducer :96-109
    bool any = false; float3 lo = photons[0].position, hi = photons[0].position; uint numValid = 0;
    for (uint i = 0; i < n; i++) {
        if (fmaxf3(photons[i].power) > 0) {
            const float3& p = photons[i].position;
            if (!any) { lo = hi = p; any = true; }
            else { lo = fminf3(lo, p); hi = fmaxf3(hi, p); }
            numValid++;
        }
    }
    // padAABB :135-141
    float3 first = lo - 0.0000001f, second = hi + 0.0000001f;
    float3 sceneWorldOrigo = first;
    float3 sceneExtent = second - first;

    float smallestPossibleCellSize = getSmallestPossibleCellSize(sceneExtent, PHOTON_GRID_MAX_SIZE_);
    float cellSize = (float)(smallestPossibleCellSize + 0.001);   // float + double literal (:230)
    uint3 gridSize = calculateGridSize(sceneExtent, cellSize);
    uint numHashCells = gridSize.x * gridSize.y * gridSize.z;
    g.cellSize = cellSize; g.gridSize = gridSize; g.origo = sceneWorldOrigo; g.numCells = numHashCells; g.numValid = numValid;
    if (numHashCells > PHOTON_GRID_MAX_SIZE_) return false;        // :244-248

    // calculateHashCellsKernel :152-173 with getPhotonGridIndex / getPhotonGridIndex1D (ppm/PhotonGrid.h:19-28)
    const uint invalidHashCell = numHashCells + 1;
    std::vector<uint> hist(numHashCells + 1, 0u);
    keys.assign(n, 0u);
    for (uint i = 0; i < n; i++) {
        uint hashCell;
        if (fmaxf3(photons[i].power) > 0) {
            float3 positivePosition = photons[i].position - sceneWorldOrigo;
            float3 sc = positivePosition * (1.f / cellSize);
            uint3 gp = {f2u_sat(floorf(sc.x)), f2u_sat(floorf(sc.y)), f2u_sat(floorf(sc.z))};
            hashCell = gp.x + gp.y * gridSize.x + gp.z * gridSize.x * gridSize.y;
            if (hashCell < numHashCells) hist[hashCell]++;   // the reference would write out of bounds here
        } else {
            hashCell = invalidHashCell;
        }
        keys[i] = hashCell;
    }
    // sortPhotonsByHash :193-196
    std::vector<uint> order(n); std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint a, uint b) { return keys[a] < keys[b]; });
    std::vector<Photon> sp(n); std::vector<uint> sk(n);
    for (uint i = 0; i < n; i++) { sp[i] = photons[order[i]]; sk[i] = keys[order[i]]; }
    std::copy(sp.begin(), sp.end(), photons);
    keys.swap(sk);
    // createHashmapOffsetTable :202-207 — exclusive scan over numHashCells+1 entries
    offsets.assign(numHashCells + 1, 0u);
    uint acc = 0;
    for (uint c = 0; c <= numHashCells; c++) { offsets[c] = acc; acc += hist[c]; }
    return true;
}

} // namespace ora
