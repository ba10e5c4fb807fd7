// LinkedCellList binning and permutation for sm_100a.
//
// Replaces the device work of
//   LinkedCellList::build            core/src/Cabana_LinkedCellList.hpp:651-739
//   LinkedCellList::storeParticleBins                                :778-823
//   permute(LinkedCellList&, data)                                   :1130-1145
//   permute(BinningData, slice)      core/src/Cabana_Sort.hpp:600-656
//
// Kernels (all HBM-bound integer/byte work; no tensor cores):
//   k_cell_count   one pass over the particles: exact locatePoint, then a
//                  warp-aggregated atomic on the cell counter that BOTH builds the
//                  histogram and claims the particle's slot in its cell (the reference
//                  needs a second atomic pass, :717-731)
//   exclusive scan decoupled look-back (cb_scan.cu)
//   k_permute_fill permute[offsets[cell] + slot] = particle id
//   k_gather_field / k_copy_back   coalesced-write gather-permute of any member
#include <new>

#include "cb_common.cuh"
#include "cb_internal.h"

namespace cb
{

namespace
{

constexpr int kBlock = 256;

__global__ void __launch_bounds__( kBlock )
    k_cell_count( PosAccess x, Grid g, int m, long long begin, long long np,
                  int* __restrict__ counts, int* __restrict__ cell_of,
                  unsigned* __restrict__ slot )
{
    const unsigned lane = lane_id();
    const unsigned lt = lanemask_lt();
    for ( long long q0 = (long long)blockIdx.x * kBlock; q0 < np;
          q0 += (long long)gridDim.x * kBlock )
    {
        const long long q = q0 + threadIdx.x;
        const bool valid = q < np;
        int c = -1 - (int)lane; // unique dummy key for idle lanes
        if ( valid )
        {
            const long long p = begin + q;
            const long long off = x.offset( p );
            const double px = x.base[off];
            const double py = x.base[off + x.comp_stride];
            const double pz = x.base[off + 2 * x.comp_stride];
            int ci = locate_1d( g, 0, px );
            int cj = locate_1d( g, 1, py );
            int ck = locate_1d( g, 2, pz );
            // Points outside [min,max] are undefined behaviour in the reference
            // (out-of-bounds write); clamp so the library never corrupts memory.
            ci = min( max( ci, 0 ), g.nx[0] - 1 );
            cj = min( max( cj, 0 ), g.nx[1] - 1 );
            ck = min( max( ck, 0 ), g.nx[2] - 1 );
            if ( m > 1 )
            {
                // Refined internal grid (Verlet build): every user cell is split into
                // m^3 sub-cells.  The USER cell stays exact (locatePoint above); the
                // sub-cell only has to be deterministic and monotone in the coordinate.
                ci = fine_index_1d( g, 0, px, ci, m );
                cj = fine_index_1d( g, 1, py, cj, m );
                ck = fine_index_1d( g, 2, pz, ck, m );
                c = ( ci * ( g.nx[1] * m ) + cj ) * ( g.nx[2] * m ) + ck;
            }
            else
                c = cardinal_index( g, ci, cj, ck );
            cell_of[q] = c;
        }
        // Warp-aggregated atomic: lanes sharing a cell issue ONE atomicAdd.
        const unsigned peers = __match_any_sync( kFullMask, c );
        const int leader = __ffs( peers ) - 1;
        int base = 0;
        if ( valid && (int)lane == leader )
            base = atomicAdd( &counts[c], __popc( peers ) );
        base = __shfl_sync( peers, base, leader );
        if ( valid )
            slot[q] = (unsigned)( base + __popc( peers & lt ) );
    }
}

__global__ void __launch_bounds__( kBlock )
    k_permute_fill( long long begin, long long np, const int* __restrict__ cell_of,
                    const unsigned* __restrict__ slot,
                    const unsigned* __restrict__ offsets,
                    unsigned* __restrict__ permute )
{
    for ( long long q = (long long)blockIdx.x * kBlock + threadIdx.x; q < np;
          q += (long long)gridDim.x * kBlock )
        permute[offsets[cell_of[q]] + slot[q]] = (unsigned)( begin + q );
}

// scratch[i*num_comp + n] = field(permute[i], n)     (Cabana_Sort.hpp:628-640)
template <class T>
__global__ void __launch_bounds__( kBlock )
    k_gather_field( FieldAccess f, long long np, const unsigned* __restrict__ permute,
                    T* __restrict__ scratch )
{
    const T* base = reinterpret_cast<const T*>( f.base );
    for ( long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < np;
          i += (long long)gridDim.x * kBlock )
    {
        const long long off = f.offset( (long long)permute[i] );
        for ( int n = 0; n < f.num_comp; ++n )
            scratch[(long long)n * np + i] = base[off + f.comp_stride * n];
    }
}

// field(begin + i, n) = scratch[...]                  (Cabana_Sort.hpp:642-655)
template <class T>
__global__ void __launch_bounds__( kBlock )
    k_copy_back( FieldAccess f, long long begin, long long np,
                 const T* __restrict__ scratch )
{
    T* base = reinterpret_cast<T*>( f.base );
    for ( long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < np;
          i += (long long)gridDim.x * kBlock )
    {
        const long long off = f.offset( begin + i );
        for ( int n = 0; n < f.num_comp; ++n )
            base[off + f.comp_stride * n] = scratch[(long long)n * np + i];
    }
}

// storeParticleBins with _sorted == true (:814-817): bins[s] = cell of sorted slot s.
__global__ void __launch_bounds__( kBlock )
    k_sorted_bins( long long begin, long long np, const unsigned* __restrict__ permute,
                   const int* __restrict__ cell_of, int* __restrict__ sorted_bins )
{
    for ( long long s = (long long)blockIdx.x * kBlock + threadIdx.x; s < np;
          s += (long long)gridDim.x * kBlock )
        sorted_bins[s] = cell_of[(long long)permute[s] - begin];
}

} // namespace

int bin_particles( const cb_grid& grid, const cb_positions& x, long long begin,
                   long long end, int* counts, unsigned* offsets, unsigned* permute,
                   int* cell_of, DeviceBuffer& rank_scratch, DeviceBuffer& scan_scratch,
                   cudaStream_t stream, int refine )
{
    const long long np = end - begin;
    const long long m3 = (long long)refine * refine * refine;
    const long long ncell = (long long)grid.nx[0] * grid.nx[1] * grid.nx[2] * m3;
    CB_CUDA( cudaMemsetAsync( counts, 0, sizeof( int ) * (size_t)ncell, stream ) );
    CB_TRY( rank_scratch.ensure( sizeof( unsigned ) * (size_t)( np > 0 ? np : 1 ), 1.1 ) );
    if ( np > 0 )
    {
        k_cell_count<<<launch_grid_for( np, kBlock ), kBlock, 0, stream>>>(
            make_access( x ), to_grid( grid ), refine, begin, np, counts, cell_of,
            rank_scratch.as<unsigned>() );
        CB_CHECK_LAUNCH();
    }
    CB_TRY( exclusive_scan_i32( counts, reinterpret_cast<int*>( offsets ), ncell, true,
                                nullptr, scan_scratch, stream ) );
    if ( np > 0 )
    {
        k_permute_fill<<<launch_grid_for( np, kBlock ), kBlock, 0, stream>>>(
            begin, np, cell_of, rank_scratch.as<unsigned>(), offsets, permute );
        CB_CHECK_LAUNCH();
    }
    return CB_OK;
}

} // namespace cb

// =====================================================================================
// C ABI
// =====================================================================================
using namespace cb;

extern "C" int cb_lcl_create( cb_lcl** out, const double* delta_h, const double* min_h,
                              const double* max_h, double neighborhood_radius,
                              double cell_size_ratio )
{
    if ( !out || !delta_h || !min_h || !max_h )
        return fail( CB_ERR_INVALID, "cb_lcl_create: null argument" );
    for ( int d = 0; d < 3; ++d )
        if ( !( delta_h[d] > 0.0 ) || !( max_h[d] > min_h[d] ) )
            return fail( CB_ERR_INVALID, "cb_lcl_create: bad grid" );
    cb_lcl* l = new ( std::nothrow ) cb_lcl();
    if ( !l )
        return fail( CB_ERR_NOMEM, "cb_lcl_create" );
    cb_grid_init( &l->grid, min_h, max_h, delta_h );
    // Stencil: explicit (radius, ratio) constructors :249-277; the simple ones use
    // _cell_stencil( grid_delta[0], 1.0, ... ) :186, :230.
    double radius = neighborhood_radius;
    double ratio = cell_size_ratio;
    if ( !( radius > 0.0 ) )
    {
        radius = delta_h[0];
        ratio = 1.0;
    }
    const double sdx = radius * ratio; // Cabana_LinkedCellList.hpp:60
    const double sdelta[3] = { sdx, sdx, sdx };
    cb_grid_init( &l->stencil_grid, min_h, max_h, sdelta );
    l->cell_range = cb_stencil_cell_range( ratio );
    l->num_cells = (long long)l->grid.nx[0] * l->grid.nx[1] * l->grid.nx[2];
    if ( l->num_cells <= 0 || l->num_cells >= 2147483647ll )
    {
        delete l;
        return fail( CB_ERR_INVALID, "cb_lcl_create: cell count out of int range" );
    }
    *out = l;
    return CB_OK;
}

extern "C" int cb_lcl_build( cb_lcl* l, const cb_positions* x, int64_t begin,
                             int64_t end, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::LinkedCellList::build" );
    if ( !l || !x )
        return fail( CB_ERR_INVALID, "cb_lcl_build: null argument" );
    // asserts of the reference (:659-660)
    if ( end < begin || end > x->n || begin < 0 )
        return fail( CB_ERR_INVALID, "cb_lcl_build: bad particle range" );
    if ( x->vlen < 1 )
        return fail( CB_ERR_INVALID, "cb_lcl_build: vlen < 1" );
    if ( x->n >= 4294967295ll )
        return fail( CB_ERR_UNSUPPORTED, "cb_lcl_build: more than 2^32-1 particles" );
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long np = end - begin;
    const size_t npa = (size_t)( np > 0 ? np : 1 );
    CB_TRY( l->counts.ensure( sizeof( int ) * (size_t)l->num_cells ) );
    CB_TRY( l->offsets.ensure( sizeof( unsigned ) * (size_t)( l->num_cells + 1 ) ) );
    CB_TRY( l->permute.ensure( sizeof( unsigned ) * npa, 1.1 ) );
    CB_TRY( l->bins.ensure( sizeof( int ) * npa, 1.1 ) );
    CB_TRY( bin_particles( l->grid, *x, begin, end, l->counts.as<int>(),
                           l->offsets.as<unsigned>(), l->permute.as<unsigned>(),
                           l->bins.as<int>(), l->rank, l->scan, stream ) );
    l->begin = begin;
    l->end = end;
    l->sorted = 0; // a rebuild describes unsorted data again (constructor: _sorted(false))
    l->built = true;
    return CB_OK;
}

extern "C" int cb_lcl_get( const cb_lcl* l, cb_lcl_view* v )
{
    if ( !l || !v )
        return fail( CB_ERR_INVALID, "cb_lcl_get: null argument" );
    v->grid = l->grid;
    v->stencil_grid = l->stencil_grid;
    v->cell_range = l->cell_range;
    v->sorted = l->sorted;
    v->begin = l->begin;
    v->end = l->end;
    v->num_cells = l->num_cells;
    v->counts = l->counts.as<int32_t>();
    v->offsets = l->offsets.as<uint32_t>();
    v->permute = l->permute.as<uint32_t>();
    v->particle_bins = l->bins.as<int32_t>();
    return CB_OK;
}

// permute(BinningData, data) (Cabana_Sort.hpp:600-656): tmp[i] = data[perm[i]], then
// data[begin+i] = tmp[i], one field at a time through `scratch`.
static int permute_fields( long long begin, long long end, const unsigned* perm,
                           const cb_field* fields, int num_fields, cb::DeviceBuffer& scratch,
                           cudaStream_t stream, const char* who )
{
    const long long np = end - begin;
    const int grid = launch_grid_for( np, kBlock );
    for ( int fi = 0; fi < num_fields && np > 0; ++fi )
    {
        const cb_field& f = fields[fi];
        if ( f.n < end || f.vlen < 1 || f.num_comp < 1 ||
             ( f.elem_bytes != 4 && f.elem_bytes != 8 ) )
            return fail( CB_ERR_INVALID, who );
        CB_TRY( scratch.ensure( (size_t)np * f.num_comp * f.elem_bytes, 1.1 ) );
        FieldAccess a = make_access( f );
        if ( f.elem_bytes == 8 )
        {
            k_gather_field<unsigned long long><<<grid, kBlock, 0, stream>>>(
                a, np, perm, scratch.as<unsigned long long>() );
            CB_CHECK_LAUNCH();
            k_copy_back<unsigned long long><<<grid, kBlock, 0, stream>>>(
                a, begin, np, scratch.as<unsigned long long>() );
            CB_CHECK_LAUNCH();
        }
        else
        {
            k_gather_field<unsigned><<<grid, kBlock, 0, stream>>>( a, np, perm,
                                                                   scratch.as<unsigned>() );
            CB_CHECK_LAUNCH();
            k_copy_back<unsigned><<<grid, kBlock, 0, stream>>>( a, begin, np,
                                                                scratch.as<unsigned>() );
            CB_CHECK_LAUNCH();
        }
    }
    return CB_OK;
}

extern "C" int cb_binning_permute( int64_t begin, int64_t end, const uint32_t* permute,
                                   const cb_field* fields, int num_fields,
                                   cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::permute" );
    if ( begin < 0 || end < begin || ( end > begin && !permute ) ||
         ( num_fields > 0 && !fields ) )
        return fail( CB_ERR_INVALID, "cb_binning_permute: bad argument" );
    static cb::DeviceBuffer scratch_of[cb::kMaxDevices]; // single caller per GPU (cabana_b200.h)
    cb::DeviceBuffer& scratch = scratch_of[cb::current_device_slot()];
    return permute_fields( begin, end, permute, fields, num_fields, scratch,
                           (cudaStream_t)stream_, "cb_binning_permute: bad field descriptor" );
}

extern "C" int cb_lcl_permute( cb_lcl* l, const cb_field* fields, int num_fields,
                               cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::permute" );
    if ( !l || ( num_fields > 0 && !fields ) )
        return fail( CB_ERR_INVALID, "cb_lcl_permute: null argument" );
    if ( !l->built )
        return fail( CB_ERR_INVALID, "cb_lcl_permute: list not built" );
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long np = l->end - l->begin;
    const int grid = launch_grid_for( np, kBlock );
    CB_TRY( permute_fields( l->begin, l->end, l->permute.as<unsigned>(), fields, num_fields,
                            l->field_scratch, stream, "cb_lcl_permute: bad field descriptor" ) );
    // update(true) + storeParticleBins() (:1141-1144)
    if ( !l->sorted && np > 0 )
    {
        CB_TRY( l->bins_alt.ensure( sizeof( int ) * (size_t)np, 1.1 ) );
        k_sorted_bins<<<grid, kBlock, 0, stream>>>( l->begin, np,
                                                    l->permute.as<unsigned>(),
                                                    l->bins.as<int>(),
                                                    l->bins_alt.as<int>() );
        CB_CHECK_LAUNCH();
        std::swap( l->bins.ptr, l->bins_alt.ptr );
        std::swap( l->bins.capacity, l->bins_alt.capacity );
    }
    l->sorted = 1;
    return CB_OK;
}

extern "C" int cb_lcl_update( cb_lcl* l, int sorted )
{
    if ( !l )
        return fail( CB_ERR_INVALID, "cb_lcl_update: null argument" );
    l->sorted = sorted ? 1 : 0;
    return CB_OK;
}

extern "C" int cb_lcl_destroy( cb_lcl* l )
{
    delete l;
    return CB_OK;
}
