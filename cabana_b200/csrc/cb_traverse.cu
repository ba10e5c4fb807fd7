// neighbor_parallel_for / neighbor_parallel_reduce over a VerletList (FirstNeighborsTag).
//
// Replaces the device loops of core/src/Cabana_Parallel.hpp:
//   :280-288  Serial  -- RangePolicy over i, thread-local loop over neighbours
//   :416-430  Team    -- TeamPolicy league = particles, TeamThreadRange over neighbours
//   :670-684 / :820-843  the same with a scalar reduction
// through the NeighborList<VerletList<..CSR|2D..>> accessors
// (core/src/Cabana_VerletList.hpp:1633-1648, :1683-1697).
//
// Serial = one thread per particle; Team = one warp per particle with lanes striding the
// row (coalesced id loads) and a shuffle reduction.  The consumers compiled here are the
// Lennard-Jones functor (north_star) and the reference unit tests' id-sum functor; generic
// user functors are instantiated by the header-only C++ shim with the same loop shapes.
#include "cb_common.cuh"
#include "cb_internal.h"

namespace cb
{
namespace
{

constexpr int kBlock = 256;

struct ListAccess
{
    const int* counts;
    const int* offsets; // null for 2D
    const int* neighbors;
    long long row_stride;
    long long col_stride;

    CB_D int num( long long i ) const { return counts[i]; }
    CB_D long long row( long long i ) const
    {
        return offsets ? (long long)offsets[i] : i * row_stride;
    }
    CB_D long long step() const { return offsets ? 1 : col_stride; }
};

ListAccess make_list( const cb_verlet_view& v )
{
    ListAccess l;
    l.counts = v.counts;
    l.offsets = v.layout == CB_LAYOUT_CSR ? v.offsets : nullptr;
    l.neighbors = v.neighbors;
    l.row_stride = v.row_stride;
    l.col_stride = v.col_stride;
    return l;
}

struct LJ
{
    double rc2;
    double s2;
    double eps24;
    double eps4;
};

CB_D void lj_pair( const LJ& p, double dx, double dy, double dz, double& fx, double& fy,
                   double& fz, bool& within )
{
    const double r2 = dx * dx + dy * dy + dz * dz;
    within = r2 < p.rc2;
    if ( within )
    {
        // one FP64 reciprocal instead of two divisions (the kernel is bound by the FP64
        // pipe); differs from s2/r2, fpair/r2 by an ulp or two, far inside the 1e-12 bar
        const double inv = 1.0 / r2;
        const double sr2 = p.s2 * inv;
        const double sr6 = sr2 * sr2 * sr2;
        const double fpair = p.eps24 * sr6 * ( 2.0 * sr6 - 1.0 ) * inv;
        fx = fpair * dx;
        fy = fpair * dy;
        fz = fpair * dz;
    }
}

CB_D void add3( double* base, long long off, long long cs, double fx, double fy,
                double fz, bool atomic )
{
    if ( atomic )
    {
        atomicAdd( base + off, fx );
        atomicAdd( base + off + cs, fy );
        atomicAdd( base + off + 2 * cs, fz );
    }
    else
    {
        base[off] += fx;
        base[off + cs] += fy;
        base[off + 2 * cs] += fz;
    }
}

// x_j is a random gather.  In an AoSoA slice the three components of one particle sit in
// three different 32-byte sectors; the LJ kernels therefore read a packed copy
// (x, y, z, pad: exactly one sector per particle) made by k_pack_positions right before
// the traversal (one coalesced pass, 56 B per particle).  The values are the same doubles,
// so every pair force is bit-identical to reading the slice.
struct Packed
{
    const double2* p;
    CB_D void get( long long j, double& x, double& y, double& z ) const
    {
        // one 256-bit load (sm_100 LDG.E.256): ONE sector lookup per gathered particle --
        // the traversal is bound by L1 tag lookups of the gathers, not by bytes
        [[maybe_unused]] double pad;
        asm( "ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];"
             : "=d"( x ), "=d"( y ), "=d"( z ), "=d"( pad )
             : "l"( p + 2 * j ) );
    }
};

__global__ void __launch_bounds__( kBlock )
    k_pack_positions( PosAccess x, long long n, double2* out )
{
    for ( long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n;
          i += (long long)gridDim.x * kBlock )
    {
        const long long o = x.offset( i );
        out[2 * i] = make_double2( x.base[o], x.base[o + x.comp_stride] );
        out[2 * i + 1] = make_double2( x.base[o + 2 * x.comp_stride], 0.0 );
    }
}

constexpr int kTeam = 8; // lanes per particle in the Team kernels (4 particles per warp)

CB_D double team_reduce_sum( double v )
{
#pragma unroll
    for ( int o = kTeam / 2; o > 0; o >>= 1 )
        v += __shfl_xor_sync( kFullMask, v, o );
    return v;
}

// Serial: Cabana_Parallel.hpp:280-288
template <bool NEWTON>
__global__ void __launch_bounds__( kBlock )
    k_lj_serial( ListAccess l, Packed x, FieldAccess f, LJ p, long long begin,
                 long long end )
{
    double* fb = reinterpret_cast<double*>( f.base );
    for ( long long i = begin + (long long)blockIdx.x * kBlock + threadIdx.x; i < end;
          i += (long long)gridDim.x * kBlock )
    {
        double xi, yi, zi;
        x.get( i, xi, yi, zi );
        const int nn = l.num( i );
        const long long row = l.row( i );
        const long long step = l.step();
        double ax = 0.0, ay = 0.0, az = 0.0;
        for ( int n = 0; n < nn; ++n )
        {
            const long long j = l.neighbors[row + n * step];
            double xj, yj, zj;
            x.get( j, xj, yj, zj );
            double fx, fy, fz;
            bool within;
            lj_pair( p, xi - xj, yi - yj, zi - zj, fx, fy, fz, within );
            if ( within )
            {
                ax += fx;
                ay += fy;
                az += fz;
                if ( NEWTON )
                    add3( fb, f.offset( j ), f.comp_stride, -fx, -fy, -fz, true );
            }
        }
        add3( fb, f.offset( i ), f.comp_stride, ax, ay, az, NEWTON );
    }
}

// Team: Cabana_Parallel.hpp:416-430 -- a team of kTeam lanes per particle (rows of ~50-80
// neighbours would leave most of a full warp idle and pay a 32-lane reduction each).
template <bool NEWTON>
__global__ void __launch_bounds__( kBlock )
    k_lj_team( ListAccess l, Packed x, FieldAccess f, LJ p, long long begin,
               long long end )
{
    double* fb = reinterpret_cast<double*>( f.base );
    const unsigned sub = threadIdx.x & ( kTeam - 1 );
    const long long team = ( (long long)blockIdx.x * kBlock + threadIdx.x ) / kTeam;
    const long long nteams = ( (long long)gridDim.x * kBlock ) / kTeam;
    // whole warps iterate together (the team reduction is a full-mask shuffle)
    const long long span = end - begin;
    const long long rounds = ( span + nteams - 1 ) / nteams;
    for ( long long r = 0; r < rounds; ++r )
    {
        const long long i = begin + r * nteams + team;
        const bool live = i < end;
        double ax = 0.0, ay = 0.0, az = 0.0;
        if ( live )
        {
            double xi, yi, zi;
            x.get( i, xi, yi, zi );
            const int nn = l.num( i );
            const long long row = l.row( i );
            const long long step = l.step();
            for ( int n = (int)sub; n < nn; n += kTeam )
            {
                const long long j = l.neighbors[row + n * step];
                double xj, yj, zj;
                x.get( j, xj, yj, zj );
                double fx, fy, fz;
                bool within;
                lj_pair( p, xi - xj, yi - yj, zi - zj, fx, fy, fz, within );
                if ( within )
                {
                    ax += fx;
                    ay += fy;
                    az += fz;
                    if ( NEWTON )
                        add3( fb, f.offset( j ), f.comp_stride, -fx, -fy, -fz, true );
                }
            }
        }
        ax = team_reduce_sum( ax );
        ay = team_reduce_sum( ay );
        az = team_reduce_sum( az );
        if ( live && sub == 0 )
            add3( fb, f.offset( i ), f.comp_stride, ax, ay, az, NEWTON );
    }
}

// neighbor_parallel_reduce with the LJ pair energy.  TEAM selects warp-per-particle.
template <bool TEAM>
__global__ void __launch_bounds__( kBlock )
    k_lj_energy( ListAccess l, Packed x, LJ p, double scale, long long begin,
                 long long end, double* energy )
{
    __shared__ double s_part[kBlock / 32];
    const unsigned lane = lane_id();
    double e = 0.0;
    if ( TEAM )
    {
        const long long warp = ( (long long)blockIdx.x * kBlock + threadIdx.x ) >> 5;
        const long long nwarps = ( (long long)gridDim.x * kBlock ) >> 5;
        for ( long long i = begin + warp; i < end; i += nwarps )
        {
            double xi, yi, zi;
            x.get( i, xi, yi, zi );
            const int nn = l.num( i );
            const long long row = l.row( i ), step = l.step();
            for ( int n = (int)lane; n < nn; n += 32 )
            {
                double xj, yj, zj;
                x.get( l.neighbors[row + n * step], xj, yj, zj );
                const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
                const double r2 = dx * dx + dy * dy + dz * dz;
                if ( r2 < p.rc2 )
                {
                    const double sr2 = p.s2 / r2;
                    const double sr6 = sr2 * sr2 * sr2;
                    e += p.eps4 * ( sr6 * sr6 - sr6 );
                }
            }
        }
    }
    else
    {
        for ( long long i = begin + (long long)blockIdx.x * kBlock + threadIdx.x;
              i < end; i += (long long)gridDim.x * kBlock )
        {
            double xi, yi, zi;
            x.get( i, xi, yi, zi );
            const int nn = l.num( i );
            const long long row = l.row( i ), step = l.step();
            for ( int n = 0; n < nn; ++n )
            {
                double xj, yj, zj;
                x.get( l.neighbors[row + n * step], xj, yj, zj );
                const double dx = xi - xj, dy = yi - yj, dz = zi - zj;
                const double r2 = dx * dx + dy * dy + dz * dz;
                if ( r2 < p.rc2 )
                {
                    const double sr2 = p.s2 / r2;
                    const double sr6 = sr2 * sr2 * sr2;
                    e += p.eps4 * ( sr6 * sr6 - sr6 );
                }
            }
        }
    }
    e = warp_reduce_sum( e );
    if ( lane == 0 )
        s_part[threadIdx.x >> 5] = e;
    __syncthreads();
    if ( threadIdx.x < 32 )
    {
        e = ( lane < kBlock / 32 ) ? s_part[lane] : 0.0;
        e = warp_reduce_sum( e );
        if ( lane == 0 )
            atomicAdd( energy, scale * e );
    }
}

// Reference unit-test functor: result[i] += j (neighbor_unit_test.hpp:291-348).
template <bool TEAM>
__global__ void __launch_bounds__( kBlock )
    k_id_sum( ListAccess l, long long begin, long long end, long long* result )
{
    if ( TEAM )
    {
        const unsigned lane = lane_id();
        const long long warp = ( (long long)blockIdx.x * kBlock + threadIdx.x ) >> 5;
        const long long nwarps = ( (long long)gridDim.x * kBlock ) >> 5;
        for ( long long i = begin + warp; i < end; i += nwarps )
        {
            const int nn = l.num( i );
            const long long row = l.row( i ), step = l.step();
            long long acc = 0;
            for ( int n = (int)lane; n < nn; n += 32 )
                acc += l.neighbors[row + n * step];
            acc = warp_reduce_sum( acc );
            if ( lane == 0 )
                result[i] += acc;
        }
    }
    else
    {
        for ( long long i = begin + (long long)blockIdx.x * kBlock + threadIdx.x;
              i < end; i += (long long)gridDim.x * kBlock )
        {
            const int nn = l.num( i );
            const long long row = l.row( i ), step = l.step();
            long long acc = 0;
            for ( int n = 0; n < nn; ++n )
                acc += l.neighbors[row + n * step];
            result[i] += acc;
        }
    }
}


// ---------------------------------------------------------------------------------------
// neighbor_parallel_for directly on a LinkedCellList (LinkedCellParallelFor,
// core/src/Cabana_Parallel.hpp:1122-1290).  Serial: one thread per particle walks the
// stencil cells serially (:1156-1207); Team: one warp per particle, lanes stride the
// particles of each cell (:1209-1268).
// ---------------------------------------------------------------------------------------
struct LclAccess
{
    Grid grid;         // _grid: binOffset / binSize indexing
    int snx[3];        // _cell_stencil.grid._nx: decode + clip (Cabana_LinkedCellList.hpp:105-119)
    int cell_range;
    int sorted;
    long long begin;
    const int* counts;
    const unsigned* offsets;
    const unsigned* permute;
    const int* bins;
};

template <bool TEAM, int MODE> // MODE 0: count within cutoff, 1: Lennard-Jones force
__global__ void __launch_bounds__( kBlock )
    k_lcl_for( LclAccess l, PosAccess x, FieldAccess f, LJ p, double c2, int* result,
               long long begin, long long end )
{
    const unsigned lane = lane_id();
    long long first, stride;
    if ( TEAM )
    {
        first = begin + ( ( (long long)blockIdx.x * kBlock + threadIdx.x ) >> 5 );
        stride = ( (long long)gridDim.x * kBlock ) >> 5;
    }
    else
    {
        first = begin + (long long)blockIdx.x * kBlock + threadIdx.x;
        stride = (long long)gridDim.x * kBlock;
    }
    for ( long long i = first; i < end; i += stride )
    {
        const int cell = l.bins[i - l.begin];
        // getStencilCells: ijk decoded on the stencil grid, +-cell_range, clipped
        const int ci = cell / ( l.snx[1] * l.snx[2] );
        const int cj = ( cell / l.snx[2] ) % l.snx[1];
        const int ck = cell % l.snx[2];
        const int imin = max( ci - l.cell_range, 0 ), imax = min( ci + l.cell_range + 1, l.snx[0] );
        const int jmin = max( cj - l.cell_range, 0 ), jmax = min( cj + l.cell_range + 1, l.snx[1] );
        const int kmin = max( ck - l.cell_range, 0 ), kmax = min( ck + l.cell_range + 1, l.snx[2] );
        const long long xo = x.offset( i );
        const double xi = x.base[xo], yi = x.base[xo + x.comp_stride],
                     zi = x.base[xo + 2 * x.comp_stride];
        double ax = 0.0, ay = 0.0, az = 0.0;
        int cnt = 0;
        for ( int gi = imin; gi < imax; ++gi )
            for ( int gj = jmin; gj < jmax; ++gj )
            {
                // cells (gi,gj,kmin..kmax-1) are consecutive in the binned order
                const int c0 = cardinal_index( l.grid, gi, gj, kmin );
                const unsigned n0 = l.offsets[c0];
                const unsigned n1 = l.offsets[c0 + ( kmax - kmin )];
                for ( unsigned n = n0 + ( TEAM ? lane : 0u ); n < n1; n += ( TEAM ? 32u : 1u ) )
                {
                    // getParticle (:863-872)
                    const long long j = l.sorted ? (long long)n + l.begin : (long long)l.permute[n];
                    if ( j == i )
                        continue; // NeighborDiscriminator<SelfNeighborTag>
                    const long long jo = x.offset( j );
                    const double dx = xi - x.base[jo], dy = yi - x.base[jo + x.comp_stride],
                                 dz = zi - x.base[jo + 2 * x.comp_stride];
                    if ( MODE == 0 )
                    {
                        // un-contracted, like the functor of the reference test
                        const double r2 = CB_ADD( CB_ADD( CB_MUL( dx, dx ), CB_MUL( dy, dy ) ),
                                                  CB_MUL( dz, dz ) );
                        cnt += r2 <= c2 ? 1 : 0;
                    }
                    else
                    {
                        double fx, fy, fz;
                        bool within;
                        lj_pair( p, dx, dy, dz, fx, fy, fz, within );
                        if ( within )
                        {
                            ax += fx;
                            ay += fy;
                            az += fz;
                        }
                    }
                }
            }
        if ( MODE == 0 )
        {
            if ( TEAM )
                cnt = warp_reduce_sum( cnt );
            if ( !TEAM || lane == 0 )
                result[i] += cnt;
        }
        else
        {
            if ( TEAM )
            {
                ax = warp_reduce_sum( ax );
                ay = warp_reduce_sum( ay );
                az = warp_reduce_sum( az );
            }
            if ( !TEAM || lane == 0 )
                add3( reinterpret_cast<double*>( f.base ), f.offset( i ), f.comp_stride, ax, ay,
                      az, false );
        }
    }
}

int lcl_access( const cb_lcl* lcl, LclAccess& l )
{
    if ( !lcl || !lcl->built )
        return fail( CB_ERR_INVALID, "LinkedCellList not built" );
    // the offsets array is indexed with _grid cardinals over runs of kmax-kmin cells: the
    // stencil grid and the binning grid must have the same shape for that (they do for
    // every constructor that passes delta = radius*ratio, e.g. VerletList's and the
    // benchmark's; the reference would mis-index otherwise as well)
    for ( int d = 0; d < 3; ++d )
        if ( lcl->grid.nx[d] != lcl->stencil_grid.nx[d] )
            return fail( CB_ERR_UNSUPPORTED,
                         "LCL traversal: stencil grid and binning grid differ in shape" );
    l.grid = to_grid( lcl->grid );
    for ( int d = 0; d < 3; ++d )
        l.snx[d] = lcl->stencil_grid.nx[d];
    l.cell_range = lcl->cell_range;
    l.sorted = lcl->sorted;
    l.begin = lcl->begin;
    l.counts = lcl->counts.as<int>();
    l.offsets = lcl->offsets.as<unsigned>();
    l.permute = lcl->permute.as<unsigned>();
    l.bins = lcl->bins.as<int>();
    return CB_OK;
}

int check_list( const cb_verlet_view* v, int64_t begin, int64_t end, const char* who )
{
    if ( !v || !v->counts || ( v->total > 0 && !v->neighbors ) )
        return fail( CB_ERR_INVALID, who );
    if ( v->layout == CB_LAYOUT_CSR && !v->offsets )
        return fail( CB_ERR_INVALID, who );
    if ( begin < 0 || end < begin || end > v->n )
        return fail( CB_ERR_INVALID, who );
    return CB_OK;
}

int team_grid( long long items )
{
    long long blocks = ( items * 32 + kBlock - 1 ) / kBlock;
    const long long cap = (long long)kNumSMs * 64;
    if ( blocks > cap )
        blocks = cap;
    return (int)( blocks < 1 ? 1 : blocks );
}

} // namespace
} // namespace cb

using namespace cb;

namespace
{
// packed positions scratch (grow-only, one per device: the library is single-caller per GPU)
cb::DeviceBuffer& packed_scratch()
{
    static cb::DeviceBuffer b[cb::kMaxDevices];
    return b[cb::current_device_slot()];
}

int pack_positions( const cb_positions* x, cudaStream_t stream, Packed& out )
{
    cb::DeviceBuffer& b = packed_scratch();
    const long long n = x->n > 0 ? x->n : 1;
    CB_TRY( b.ensure( (size_t)n * 32, 1.1 ) );
    if ( x->n > 0 )
    {
        k_pack_positions<<<launch_grid_for( x->n, kBlock ), kBlock, 0, stream>>>(
            make_access( *x ), x->n, b.as<double2>() );
        CB_CHECK_LAUNCH();
    }
    out.p = b.as<double2>();
    return CB_OK;
}
} // namespace

extern "C" int cb_neighbor_for_lj( const cb_verlet_view* list, const cb_positions* x,
                                   const cb_field* f, double eps, double sigma,
                                   double rc, int newton, int op, int64_t begin,
                                   int64_t end, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::neighbor_parallel_for" );
    CB_TRY( check_list( list, begin, end, "cb_neighbor_for_lj: bad list or range" ) );
    if ( !x || !f || x->vlen < 1 || f->vlen < 1 )
        return fail( CB_ERR_INVALID, "cb_neighbor_for_lj: null argument" );
    if ( f->elem_bytes != 8 || f->num_comp != 3 || f->n < list->n || x->n < list->n )
        return fail( CB_ERR_INVALID, "cb_neighbor_for_lj: force field must be double[3]" );
    if ( op != CB_OP_SERIAL && op != CB_OP_TEAM )
        return fail( CB_ERR_INVALID, "cb_neighbor_for_lj: op must be Serial or Team" );
    if ( end == begin )
        return CB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    LJ p;
    p.rc2 = rc * rc;
    p.s2 = sigma * sigma;
    p.eps24 = 24.0 * eps;
    p.eps4 = 4.0 * eps;
    const ListAccess l = make_list( *list );
    Packed xa;
    CB_TRY( pack_positions( x, stream, xa ) );
    const FieldAccess fa = make_access( *f );
    const long long items = end - begin;
    if ( op == CB_OP_SERIAL )
    {
        const int grid = launch_grid_for( items, kBlock );
        if ( newton )
            k_lj_serial<true><<<grid, kBlock, 0, stream>>>( l, xa, fa, p, begin, end );
        else
            k_lj_serial<false><<<grid, kBlock, 0, stream>>>( l, xa, fa, p, begin, end );
    }
    else
    {
        const int grid = launch_grid_for( items * kTeam, kBlock );
        if ( newton )
            k_lj_team<true><<<grid, kBlock, 0, stream>>>( l, xa, fa, p, begin, end );
        else
            k_lj_team<false><<<grid, kBlock, 0, stream>>>( l, xa, fa, p, begin, end );
    }
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_neighbor_reduce_lj( const cb_verlet_view* list, const cb_positions* x,
                                      double eps, double sigma, double rc, double scale,
                                      int op, int64_t begin, int64_t end,
                                      double* energy_h, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::neighbor_parallel_reduce" );
    CB_TRY( check_list( list, begin, end, "cb_neighbor_reduce_lj: bad list or range" ) );
    if ( !x || !energy_h || x->vlen < 1 || x->n < list->n )
        return fail( CB_ERR_INVALID, "cb_neighbor_reduce_lj: null argument" );
    if ( op != CB_OP_SERIAL && op != CB_OP_TEAM )
        return fail( CB_ERR_INVALID, "cb_neighbor_reduce_lj: op must be Serial or Team" );
    cudaStream_t stream = (cudaStream_t)stream_;
    double* energy_dev = nullptr;
    CB_CUDA( cudaMallocAsync( (void**)&energy_dev, sizeof( double ), stream ) );
    CB_CUDA( cudaMemsetAsync( energy_dev, 0, sizeof( double ), stream ) );
    LJ p;
    p.rc2 = rc * rc;
    p.s2 = sigma * sigma;
    p.eps24 = 24.0 * eps;
    p.eps4 = 4.0 * eps;
    const long long items = end - begin;
    if ( items > 0 )
    {
        Packed xa;
        CB_TRY( pack_positions( x, stream, xa ) );
        if ( op == CB_OP_SERIAL )
            k_lj_energy<false><<<launch_grid_for( items, kBlock ), kBlock, 0, stream>>>(
                make_list( *list ), xa, p, scale, begin, end, energy_dev );
        else
            k_lj_energy<true><<<team_grid( items ), kBlock, 0, stream>>>(
                make_list( *list ), xa, p, scale, begin, end, energy_dev );
        CB_CHECK_LAUNCH();
    }
    CB_CUDA( cudaMemcpyAsync( energy_h, energy_dev, sizeof( double ),
                              cudaMemcpyDeviceToHost, stream ) );
    CB_CUDA( cudaFreeAsync( energy_dev, stream ) );
    CB_CUDA( cudaStreamSynchronize( stream ) );
    return CB_OK;
}

extern "C" int cb_neighbor_for_id_sum( const cb_verlet_view* list, int64_t* result,
                                       int op, int64_t begin, int64_t end,
                                       cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::neighbor_parallel_for" );
    CB_TRY( check_list( list, begin, end, "cb_neighbor_for_id_sum: bad list or range" ) );
    if ( !result )
        return fail( CB_ERR_INVALID, "cb_neighbor_for_id_sum: null argument" );
    if ( op != CB_OP_SERIAL && op != CB_OP_TEAM )
        return fail( CB_ERR_INVALID, "cb_neighbor_for_id_sum: op must be Serial or Team" );
    if ( end == begin )
        return CB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long items = end - begin;
    if ( op == CB_OP_SERIAL )
        k_id_sum<false><<<launch_grid_for( items, kBlock ), kBlock, 0, stream>>>(
            make_list( *list ), begin, end, reinterpret_cast<long long*>( result ) );
    else
        k_id_sum<true><<<team_grid( items ), kBlock, 0, stream>>>(
            make_list( *list ), begin, end, reinterpret_cast<long long*>( result ) );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_lcl_neighbor_for_lj( const cb_lcl* lcl, const cb_positions* x,
                                       const cb_field* f, double eps, double sigma, double rc,
                                       int op, int64_t begin, int64_t end, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::neighbor_parallel_for" );
    LclAccess l;
    CB_TRY( lcl_access( lcl, l ) );
    if ( !x || !f || x->vlen < 1 || f->vlen < 1 || f->elem_bytes != 8 || f->num_comp != 3 )
        return fail( CB_ERR_INVALID, "cb_lcl_neighbor_for_lj: bad argument" );
    if ( begin < lcl->begin || end > lcl->end || end < begin || x->n < lcl->end || f->n < end )
        return fail( CB_ERR_INVALID, "cb_lcl_neighbor_for_lj: range outside the binned range" );
    if ( op != CB_OP_SERIAL && op != CB_OP_TEAM )
        return fail( CB_ERR_INVALID, "cb_lcl_neighbor_for_lj: op must be Serial or Team" );
    if ( end == begin )
        return CB_OK;
    LJ p;
    p.rc2 = rc * rc;
    p.s2 = sigma * sigma;
    p.eps24 = 24.0 * eps;
    p.eps4 = 4.0 * eps;
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long items = end - begin;
    if ( op == CB_OP_SERIAL )
        k_lcl_for<false, 1><<<launch_grid_for( items, kBlock ), kBlock, 0, stream>>>(
            l, make_access( *x ), make_access( *f ), p, 0.0, nullptr, begin, end );
    else
        k_lcl_for<true, 1><<<team_grid( items ), kBlock, 0, stream>>>(
            l, make_access( *x ), make_access( *f ), p, 0.0, nullptr, begin, end );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_lcl_neighbor_for_count( const cb_lcl* lcl, const cb_positions* x,
                                          double cutoff, int32_t* result, int op,
                                          int64_t begin, int64_t end, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::neighbor_parallel_for" );
    LclAccess l;
    CB_TRY( lcl_access( lcl, l ) );
    if ( !x || !result || x->vlen < 1 )
        return fail( CB_ERR_INVALID, "cb_lcl_neighbor_for_count: bad argument" );
    if ( begin < lcl->begin || end > lcl->end || end < begin || x->n < lcl->end )
        return fail( CB_ERR_INVALID, "cb_lcl_neighbor_for_count: range outside the binned range" );
    if ( op != CB_OP_SERIAL && op != CB_OP_TEAM )
        return fail( CB_ERR_INVALID, "cb_lcl_neighbor_for_count: op must be Serial or Team" );
    if ( end == begin )
        return CB_OK;
    LJ p = {};
    FieldAccess fa = {};
    cudaStream_t stream = (cudaStream_t)stream_;
    const long long items = end - begin;
    if ( op == CB_OP_SERIAL )
        k_lcl_for<false, 0><<<launch_grid_for( items, kBlock ), kBlock, 0, stream>>>(
            l, make_access( *x ), fa, p, cutoff * cutoff, result, begin, end );
    else
        k_lcl_for<true, 0><<<team_grid( items ), kBlock, 0, stream>>>(
            l, make_access( *x ), fa, p, cutoff * cutoff, result, begin, end );
    CB_CHECK_LAUNCH();
    return CB_OK;
}
