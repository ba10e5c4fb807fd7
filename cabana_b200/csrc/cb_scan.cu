// Single-pass exclusive prefix scan (decoupled look-back) and max/sum reduction.
//
// Replaces the reference's Kokkos::parallel_scan call sites on the hot path:
//   core/src/Cabana_LinkedCellList.hpp:701-712 (cell offsets) and
//   core/src/Cabana_VerletList.hpp:507-523 (CSR row offsets, OffsetScanOp :478-491),
// and the Kokkos::Max reduction of :536-551.
//
// HBM-bound: one read + one write of 4 B per item.  Each 2048-item tile publishes its
// aggregate, then its inclusive prefix, in a single 64-bit status word (2 flag bits +
// 62 value bits), so no fence is needed between flag and payload.  Tiles are handed out
// by an atomic ticket so a tile's predecessors are always resident or finished.
#include "cb_common.cuh"
#include "cb_internal.h"

namespace cb
{

namespace
{

constexpr int kScanThreads = 512;
constexpr int kScanVec = 4;                 // int4 vectors per thread
constexpr int kScanItems = 4 * kScanVec;    // 16 consecutive items per thread
constexpr int kScanTile = kScanThreads * kScanItems; // 8192 items per tile
constexpr int kScanWarps = kScanThreads / 32;

using ull = unsigned long long;
constexpr ull kValueMask = ( 1ull << 62 ) - 1ull;

CB_D ull pack_status( unsigned flag, long long v )
{
    return ( (ull)flag << 62 ) | ( (ull)v & kValueMask );
}
CB_D ull load_status( const ull* p )
{
    return *reinterpret_cast<const volatile ull*>( p );
}
CB_D void store_status( ull* p, ull v )
{
    *reinterpret_cast<volatile ull*>( p ) = v;
}

__global__ void __launch_bounds__( kScanThreads )
    k_exclusive_scan( const int* __restrict__ in, int* out, long long n,
                      int write_total_at_n, long long* total_dev, ull* status,
                      unsigned* tile_counter, unsigned num_tiles, int aligned16,
                      long long* stats /* optional: [0] = max (zeroed by the host), [1] = sum */ )
{
    __shared__ unsigned s_tile;
    __shared__ long long s_warp_excl[kScanWarps];
    __shared__ long long s_tile_prefix;

    const int t = threadIdx.x;
    const int lane = t & 31;
    const int warp = t >> 5;

    if ( t == 0 )
        s_tile = atomicAdd( tile_counter, 1u );
    __syncthreads();
    const unsigned tile = s_tile;

    const long long base = (long long)tile * kScanTile + (long long)t * kScanItems;
    int v[kScanItems];
    if ( aligned16 && base + kScanItems <= n )
    {
#pragma unroll
        for ( int k = 0; k < kScanVec; ++k )
        {
            const int4 q = *reinterpret_cast<const int4*>( in + base + 4 * k );
            v[4 * k + 0] = q.x;
            v[4 * k + 1] = q.y;
            v[4 * k + 2] = q.z;
            v[4 * k + 3] = q.w;
        }
    }
    else
    {
#pragma unroll
        for ( int e = 0; e < kScanItems; ++e )
            v[e] = ( base + e < n ) ? in[base + e] : 0;
    }
    if ( stats )
    {
        int mx = v[0];
#pragma unroll
        for ( int e = 1; e < kScanItems; ++e )
            mx = max( mx, v[e] );
        mx = warp_reduce_max( mx );
        if ( lane == 0 && mx > 0 )
            atomicMax( reinterpret_cast<long long*>( &stats[0] ), (long long)mx );
    }
    long long tsum = 0;
#pragma unroll
    for ( int e = 0; e < kScanItems; ++e )
        tsum += v[e];

    // Inclusive scan of thread sums inside the warp.
    long long incl = tsum;
#pragma unroll
    for ( int o = 1; o < 32; o <<= 1 )
    {
        long long y = __shfl_up_sync( kFullMask, incl, o );
        if ( lane >= o )
            incl += y;
    }
    if ( lane == 31 )
        s_warp_excl[warp] = incl;
    __syncthreads();

    if ( warp == 0 )
    {
        // Scan the warp totals; publish; look back.
        long long w = ( lane < kScanWarps ) ? s_warp_excl[lane] : 0;
        long long wi = w;
#pragma unroll
        for ( int o = 1; o < 32; o <<= 1 )
        {
            long long y = __shfl_up_sync( kFullMask, wi, o );
            if ( lane >= o )
                wi += y;
        }
        if ( lane < kScanWarps )
            s_warp_excl[lane] = wi - w;
        const long long agg = __shfl_sync( kFullMask, wi, 31 );

        long long prefix = 0;
        if ( tile == 0 )
        {
            if ( lane == 0 )
                store_status( &status[0], pack_status( 2u, agg ) );
        }
        else
        {
            if ( lane == 0 )
                store_status( &status[tile], pack_status( 1u, agg ) );
            long long idx = (long long)tile - 1 - lane;
            long long running = 0;
            while ( true )
            {
                ull sv;
                do
                {
                    sv = ( idx >= 0 ) ? load_status( &status[idx] )
                                      : pack_status( 2u, 0 );
                } while ( __any_sync( kFullMask, ( sv >> 62 ) == 0ull ) );
                const unsigned flag = (unsigned)( sv >> 62 );
                const long long val = (long long)( sv & kValueMask );
                const unsigned incl_mask = __ballot_sync( kFullMask, flag == 2u );
                if ( incl_mask )
                {
                    // lane 0 is the nearest predecessor: sum up to the first lane
                    // that already carries an inclusive prefix.
                    const int first = __ffs( incl_mask ) - 1;
                    running += warp_reduce_sum( ( lane <= first ) ? val : 0ll );
                    break;
                }
                running += warp_reduce_sum( val );
                idx -= 32;
            }
            prefix = running;
            if ( lane == 0 )
                store_status( &status[tile], pack_status( 2u, prefix + agg ) );
        }
        if ( lane == 0 )
        {
            s_tile_prefix = prefix;
            if ( tile == num_tiles - 1 )
            {
                if ( total_dev )
                    *total_dev = prefix + agg;
                if ( stats )
                    stats[1] = prefix + agg;
                if ( write_total_at_n )
                    out[n] = (int)( prefix + agg );
            }
        }
    }
    __syncthreads();

    long long run = s_tile_prefix + s_warp_excl[warp] + ( incl - tsum );
    if ( aligned16 && base + kScanItems <= n )
    {
#pragma unroll
        for ( int k = 0; k < kScanVec; ++k )
        {
            int4 q;
            q.x = (int)run;
            run += v[4 * k + 0];
            q.y = (int)run;
            run += v[4 * k + 1];
            q.z = (int)run;
            run += v[4 * k + 2];
            q.w = (int)run;
            run += v[4 * k + 3];
            *reinterpret_cast<int4*>( out + base + 4 * k ) = q;
        }
    }
    else
    {
#pragma unroll
        for ( int e = 0; e < kScanItems; ++e )
        {
            if ( base + e < n )
                out[base + e] = (int)run;
            run += v[e];
        }
    }
}

__global__ void k_scan_empty( int* out, int write_total_at_n, long long* total_dev )
{
    if ( total_dev )
        *total_dev = 0;
    if ( write_total_at_n )
        out[0] = 0;
}

__global__ void __launch_bounds__( 256 )
    k_max_and_sum( const int* __restrict__ in, long long n, long long* stats )
{
    __shared__ int s_max[8];
    __shared__ long long s_sum[8];
    int mx = 0;
    long long sum = 0;
    for ( long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n;
          i += (long long)gridDim.x * blockDim.x )
    {
        const int v = in[i];
        mx = max( mx, v );
        sum += v;
    }
    mx = warp_reduce_max( mx );
    sum = warp_reduce_sum( sum );
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if ( lane == 0 )
    {
        s_max[warp] = mx;
        s_sum[warp] = sum;
    }
    __syncthreads();
    if ( warp == 0 )
    {
        mx = ( lane < 8 ) ? s_max[lane] : 0;
        sum = ( lane < 8 ) ? s_sum[lane] : 0;
        mx = warp_reduce_max( mx );
        sum = warp_reduce_sum( sum );
        if ( lane == 0 )
        {
            atomicMax( reinterpret_cast<long long*>( &stats[0] ), (long long)mx );
            atomicAdd( reinterpret_cast<ull*>( &stats[1] ), (ull)sum );
        }
    }
}

} // namespace

int launch_grid_for( long long work_items, int block )
{
    // Enough CTAs to fill the machine a few times over, capped so grid-stride loops
    // amortise launch cost: multiples of the SM count (148 on B200).
    long long blocks = ( work_items + block - 1 ) / block;
    const long long cap = (long long)kNumSMs * 16;
    if ( blocks > cap )
        blocks = cap;
    if ( blocks < 1 )
        blocks = 1;
    return (int)blocks;
}

static int scan_impl( const int* in, int* out, long long n, bool write_total_at_n,
                      long long* total_dev, long long* stats_dev, DeviceBuffer& scratch,
                      cudaStream_t stream )
{
    if ( n < 0 )
        return fail( CB_ERR_INVALID, "exclusive_scan_i32: n < 0" );
    if ( stats_dev )
        CB_CUDA( cudaMemsetAsync( stats_dev, 0, 2 * sizeof( long long ), stream ) );
    if ( n == 0 )
    {
        k_scan_empty<<<1, 1, 0, stream>>>( out, write_total_at_n ? 1 : 0, total_dev );
        CB_CHECK_LAUNCH();
        return CB_OK;
    }
    const long long num_tiles = ( n + kScanTile - 1 ) / kScanTile;
    const size_t bytes = 16 + (size_t)num_tiles * sizeof( ull );
    CB_TRY( scratch.ensure( bytes, 1.25 ) );
    CB_CUDA( cudaMemsetAsync( scratch.ptr, 0, bytes, stream ) );
    unsigned* counter = scratch.as<unsigned>();
    ull* status = reinterpret_cast<ull*>( scratch.as<char>() + 16 );
    const int aligned16 = ( ( (uintptr_t)in | (uintptr_t)out ) & 15u ) == 0 ? 1 : 0;
    k_exclusive_scan<<<(unsigned)num_tiles, kScanThreads, 0, stream>>>(
        in, out, n, write_total_at_n ? 1 : 0, total_dev, status, counter,
        (unsigned)num_tiles, aligned16, stats_dev );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

int exclusive_scan_i32( const int* in, int* out, long long n, bool write_total_at_n,
                        long long* total_dev, DeviceBuffer& scratch,
                        cudaStream_t stream )
{
    return scan_impl( in, out, n, write_total_at_n, total_dev, nullptr, scratch, stream );
}

// The same scan that also leaves max(in) in stats_dev[0] and sum(in) in stats_dev[1]
// (processCounts needs both: Cabana_VerletList.hpp:507-523 and :536-551) in the one pass.
int exclusive_scan_stats_i32( const int* in, int* out, long long n, long long* stats_dev,
                              DeviceBuffer& scratch, cudaStream_t stream )
{
    return scan_impl( in, out, n, false, nullptr, stats_dev, scratch, stream );
}

int max_and_sum_i32( const int* in, long long n, long long* stats_dev,
                     cudaStream_t stream )
{
    CB_CUDA( cudaMemsetAsync( stats_dev, 0, 2 * sizeof( long long ), stream ) );
    if ( n > 0 )
    {
        k_max_and_sum<<<launch_grid_for( n, 256 * 4 ), 256, 0, stream>>>( in, n,
                                                                            stats_dev );
        CB_CHECK_LAUNCH();
    }
    return CB_OK;
}

} // namespace cb
