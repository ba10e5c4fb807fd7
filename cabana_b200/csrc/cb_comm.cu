// Slab Halo / Distributor device kernels for sm_100a.
//
// Replaces the device work of
//   countSendsAndCreateSteering   core/src/Cabana_CommunicationPlanBase.hpp:96-224
//   createExportSteering          :596-657
//   gather pack / unpack          core/src/impl/Cabana_Halo_Mpi.hpp:58-65, :113-121
//   scatter (atomic add unpack)   :269-282, :334-347
//   migrate pack / unpack         core/src/impl/Cabana_Migrate_Mpi.hpp:92-103, :164-172
//   destination-from-position     grid/src/Cabana_Grid_ParticleDistributor.hpp:119-158
//
// The reference claims steering slots with atomics, so the order inside a destination
// block is not deterministic there (core/unit_test/tstDistributor.hpp:242-244).  Here a
// stable partition (flag -> decoupled look-back scan -> scatter, one pass per non-empty
// destination) makes it deterministic: ascending destination rank, ascending export index.
#include <new>
#include <string.h>

#include <vector>

#include "cb_common.cuh"
#include "cb_internal.h"

namespace cb
{
namespace
{

constexpr int kBlock = 256;
constexpr int kMaxRanks = 1024;
constexpr int kMaxFields = 8;
constexpr int kMaxBounds = 65;

struct Bounds
{
    double b[kMaxBounds];
    int num_ranks;
};

__global__ void __launch_bounds__( kBlock )
    k_slab_halo_select( PosAccess x, long long num_local, double lo_thresh,
                        double hi_thresh, int lo_rank, int hi_rank,
                        int* __restrict__ export_ranks, unsigned* __restrict__ export_ids )
{
    for ( long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < num_local;
          i += (long long)gridDim.x * kBlock )
    {
        const double px = x.base[x.offset( i )];
        export_ranks[2 * i] = ( lo_rank >= 0 && px < lo_thresh ) ? lo_rank : -1;
        export_ranks[2 * i + 1] = ( hi_rank >= 0 && px >= hi_thresh ) ? hi_rank : -1;
        export_ids[2 * i] = (unsigned)i;
        export_ids[2 * i + 1] = (unsigned)i;
    }
}

__global__ void __launch_bounds__( kBlock )
    k_slab_destinations( PosAccess x, long long num_local, Bounds bd,
                         int* __restrict__ export_ranks )
{
    for ( long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < num_local;
          i += (long long)gridDim.x * kBlock )
    {
        const double px = x.base[x.offset( i )];
        int dest = -1;
        if ( px >= bd.b[0] && px <= bd.b[bd.num_ranks] )
        {
            dest = bd.num_ranks - 1; // the last slab owns its upper face
            for ( int g = 0; g < bd.num_ranks - 1; ++g )
                if ( px < bd.b[g + 1] )
                {
                    dest = g;
                    break;
                }
        }
        export_ranks[i] = dest;
    }
}

__global__ void __launch_bounds__( kBlock )
    k_rank_histogram( const int* __restrict__ export_ranks, long long n, int num_ranks,
                      unsigned long long* __restrict__ counts )
{
    extern __shared__ unsigned s_hist[];
    for ( int r = threadIdx.x; r < num_ranks; r += kBlock )
        s_hist[r] = 0;
    __syncthreads();
    for ( long long i = (long long)blockIdx.x * kBlock + threadIdx.x; i < n;
          i += (long long)gridDim.x * kBlock )
    {
        const int r = export_ranks[i];
        if ( r >= 0 && r < num_ranks )
            atomicAdd( &s_hist[r], 1u );
    }
    __syncthreads();
    for ( int r = threadIdx.x; r < num_ranks; r += kBlock )
        if ( s_hist[r] )
            atomicAdd( &counts[r], (unsigned long long)s_hist[r] );
}

// Stable partition of the exports by destination in ONE pass over them per 5 bits of the
// (dense) destination index -- one pass for up to 32 destinations that actually receive
// something, two up to 1024 -- instead of a flag / scan / scatter sweep per destination.
// Least-significant-digit radix sort of (dense destination, export index): a tile of 256 exports
// per CTA; inside a warp __match_any groups equal digits (rank = peers below me), per-warp digit
// counts in shared memory order the warps of a tile, and a scan of the digit-major table of tile
// histograms orders the tiles.  Dropped exports (rank -1) leave in the first pass.
constexpr int kSteerDigits = 32;
constexpr int kSteerTile = 256;

struct SteerKey
{
    bool valid;
    unsigned src;
    int digit;
};
template <bool FIRST>
CB_D SteerKey steer_key( const int* __restrict__ export_ranks,
                         const unsigned* __restrict__ idx_in, long long count,
                         const int* __restrict__ dense_of, int shift )
{
    SteerKey k;
    const long long i = (long long)blockIdx.x * kSteerTile + threadIdx.x;
    k.valid = i < count;
    k.src = 0u;
    k.digit = 0;
    if ( k.valid )
    {
        k.src = FIRST ? (unsigned)i : idx_in[i];
        const int r = export_ranks[k.src];
        const int d = r >= 0 ? dense_of[r] : -1;
        k.valid = d >= 0;
        k.digit = ( d >> shift ) & ( kSteerDigits - 1 );
    }
    return k;
}

template <bool FIRST>
__global__ void __launch_bounds__( kSteerTile )
    k_steer_digit_hist( const int* __restrict__ export_ranks,
                        const unsigned* __restrict__ idx_in, long long count,
                        const int* __restrict__ dense_of, int shift, int* __restrict__ hist,
                        long long num_tiles )
{
    __shared__ int s_cnt[kSteerDigits];
    if ( threadIdx.x < kSteerDigits )
        s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const SteerKey k = steer_key<FIRST>( export_ranks, idx_in, count, dense_of, shift );
    const unsigned lane = lane_id();
    const unsigned peers =
        __match_any_sync( kFullMask, k.valid ? k.digit : kSteerDigits + (int)lane );
    if ( k.valid && (int)lane == __ffs( (int)peers ) - 1 )
        atomicAdd( &s_cnt[k.digit], __popc( peers ) );
    __syncthreads();
    if ( threadIdx.x < kSteerDigits )
        hist[(long long)threadIdx.x * num_tiles + blockIdx.x] = s_cnt[threadIdx.x];
}

template <bool FIRST, bool LAST>
__global__ void __launch_bounds__( kSteerTile )
    k_steer_digit_scatter( const int* __restrict__ export_ranks,
                           const unsigned* __restrict__ idx_in, long long count,
                           const int* __restrict__ dense_of, int shift,
                           const int* __restrict__ base, long long num_tiles,
                           const unsigned* __restrict__ export_ids, unsigned* __restrict__ out )
{
    __shared__ int s_w[kSteerTile / 32][kSteerDigits];
    const unsigned lane = lane_id();
    const int warp = threadIdx.x >> 5;
    s_w[warp][lane] = 0;
    __syncthreads();
    const SteerKey k = steer_key<FIRST>( export_ranks, idx_in, count, dense_of, shift );
    const unsigned peers =
        __match_any_sync( kFullMask, k.valid ? k.digit : kSteerDigits + (int)lane );
    const int rank_in_warp = __popc( peers & lanemask_lt() );
    if ( k.valid && rank_in_warp == 0 )
        s_w[warp][k.digit] = __popc( peers );
    __syncthreads();
    if ( !k.valid )
        return;
    int pos = base[(long long)k.digit * num_tiles + blockIdx.x] + rank_in_warp;
    for ( int w = 0; w < warp; ++w )
        pos += s_w[w][k.digit];
    out[pos] = LAST ? ( export_ids ? export_ids[k.src] : k.src ) : k.src;
}

struct FieldSet
{
    FieldAccess f[kMaxFields];
    int byte_off[kMaxFields];
    int num;
    int tuple_bytes;
};

// DIR 0 = pack by steering (field -> buffer), 1 = unpack into a range (buffer -> field),
// 2 = pack a contiguous range [dst_begin, dst_begin + count) (field -> buffer)
template <int DIR>
__global__ void __launch_bounds__( kBlock )
    k_pack_unpack( FieldSet fs, const unsigned* __restrict__ steering,
                   long long dst_begin, long long count, char* buffer )
{
    for ( long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < count;
          j += (long long)gridDim.x * kBlock )
    {
        const long long elem = DIR == 0 ? (long long)steering[j] : dst_begin + j;
        constexpr bool kToBuffer = DIR != 1;
        char* tuple = buffer + j * fs.tuple_bytes;
        for ( int k = 0; k < fs.num; ++k )
        {
            const FieldAccess& f = fs.f[k];
            const long long off = f.offset( elem );
            if ( f.elem_bytes == 8 )
            {
                unsigned long long* fb = reinterpret_cast<unsigned long long*>( f.base );
                unsigned long long* tb =
                    reinterpret_cast<unsigned long long*>( tuple + fs.byte_off[k] );
                for ( int c = 0; c < f.num_comp; ++c )
                {
                    if ( kToBuffer )
                        tb[c] = fb[off + f.comp_stride * c];
                    else
                        fb[off + f.comp_stride * c] = tb[c];
                }
            }
            else
            {
                unsigned* fb = reinterpret_cast<unsigned*>( f.base );
                unsigned* tb = reinterpret_cast<unsigned*>( tuple + fs.byte_off[k] );
                for ( int c = 0; c < f.num_comp; ++c )
                {
                    if ( kToBuffer )
                        tb[c] = fb[off + f.comp_stride * c];
                    else
                        fb[off + f.comp_stride * c] = tb[c];
                }
            }
        }
    }
}

// Cabana::scatter sums ghost contributions into their owners for any arithmetic slice value type
// (impl/Cabana_Halo_Mpi.hpp:334-347, Kokkos::atomic_add on the slice's value_type).
// recv holds one packed tuple of this single field per export slot (the layout cb_comm_pack /
// cb_comm_pack_range write: `tuple_elems` values of T per tuple, the first num_comp used).
template <class T>
__global__ void __launch_bounds__( kBlock )
    k_scatter_add( FieldAccess f, const unsigned* __restrict__ steering, long long count,
                   const T* __restrict__ recv, int tuple_elems )
{
    T* fb = reinterpret_cast<T*>( f.base );
    for ( long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < count;
          j += (long long)gridDim.x * kBlock )
    {
        const long long off = f.offset( (long long)steering[j] );
        for ( int c = 0; c < f.num_comp; ++c )
            atomicAdd( fb + off + f.comp_stride * c, recv[j * tuple_elems + c] );
    }
}

// Fused slab-halo plan: ghost selection + STABLE compaction of the two export lists in one
// kernel (decoupled look-back over a packed (lo,hi) pair of counters), so a plan costs one
// launch and one 16-byte read-back instead of ten launches and two syncs.
constexpr int kHaloThreads = 256;
constexpr int kHaloItems = 16; // per thread: few, fat tiles keep the look-back chain short
constexpr int kHaloTile = kHaloThreads * kHaloItems;

using ull = unsigned long long;

// A warp owns 32 * kHaloItems consecutive particles; item e of lane l is particle
// warp_base + 32 e + l, so every load instruction reads 256 contiguous bytes of an AoSoA and the
// stable order inside the warp is e-major: a ballot per item gives each lane its place (no
// shuffle scans), and the ballots themselves are all a thread has to keep between the two sweeps.
__global__ void __launch_bounds__( kHaloThreads )
    k_halo_compact( PosAccess x, long long n, double lo_thresh, double hi_thresh, int has_lo,
                    int has_hi, unsigned* __restrict__ steer_lo,
                    unsigned* __restrict__ steer_hi, ull* status, unsigned* tile_counter,
                    unsigned num_tiles, long long* totals )
{
    __shared__ unsigned s_tile;
    __shared__ ull s_warp[kHaloThreads / 32];
    __shared__ ull s_prefix;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    if ( t == 0 )
        s_tile = atomicAdd( tile_counter, 1u );
    __syncthreads();
    const unsigned tile = s_tile;
    const long long wbase = (long long)tile * kHaloTile + (long long)warp * ( 32 * kHaloItems );
    double px[kHaloItems];
#pragma unroll
    for ( int e = 0; e < kHaloItems; ++e )
    {
        const long long p = wbase + 32 * e + lane;
        px[e] = p < n ? x.base[x.offset( p )] : 0.0;
    }
    unsigned blo[kHaloItems], bhi[kHaloItems];
    unsigned clo = 0, chi = 0;
#pragma unroll
    for ( int e = 0; e < kHaloItems; ++e )
    {
        const bool in = wbase + 32 * e + lane < n;
        blo[e] = __ballot_sync( kFullMask, in && has_lo && px[e] < lo_thresh );
        bhi[e] = __ballot_sync( kFullMask, in && has_hi && px[e] >= hi_thresh );
        clo += (unsigned)__popc( blo[e] );
        chi += (unsigned)__popc( bhi[e] );
    }
    // packed sums: lo in bits 0..30, hi in bits 31..61 (n < 2^31); clo / chi are warp totals
    const ull mine = (ull)clo | ( (ull)chi << 31 );
    if ( lane == 0 )
        s_warp[warp] = mine;
    __syncthreads();
    if ( warp == 0 )
    {
        const ull w = lane < kHaloThreads / 32 ? s_warp[lane] : 0ull;
        ull wi = w;
#pragma unroll
        for ( int o = 1; o < 32; o <<= 1 )
        {
            const ull y = __shfl_up_sync( kFullMask, wi, o );
            if ( lane >= o )
                wi += y;
        }
        if ( lane < kHaloThreads / 32 )
            s_warp[lane] = wi - w;
        const ull agg = __shfl_sync( kFullMask, wi, 31 );
        const ull vmask = ( 1ull << 62 ) - 1ull;
        ull prefix = 0ull;
        if ( tile == 0 )
        {
            if ( lane == 0 )
                *reinterpret_cast<volatile ull*>( &status[0] ) = ( 2ull << 62 ) | agg;
        }
        else
        {
            if ( lane == 0 )
                *reinterpret_cast<volatile ull*>( &status[tile] ) = ( 1ull << 62 ) | agg;
            long long idx = (long long)tile - 1 - lane;
            ull running = 0ull;
            while ( true )
            {
                ull sv;
                do
                {
                    sv = idx >= 0 ? *reinterpret_cast<const volatile ull*>( &status[idx] )
                                  : ( 2ull << 62 );
                } while ( __any_sync( kFullMask, ( sv >> 62 ) == 0ull ) );
                const unsigned flag = (unsigned)( sv >> 62 );
                const ull val = sv & vmask;
                const unsigned incl_mask = __ballot_sync( kFullMask, flag == 2u );
                ull contrib = val;
                if ( incl_mask )
                {
                    const int first = __ffs( incl_mask ) - 1;
                    contrib = lane <= first ? val : 0ull;
                }
#pragma unroll
                for ( int o = 16; o > 0; o >>= 1 )
                    contrib += __shfl_xor_sync( kFullMask, contrib, o );
                running += contrib;
                if ( incl_mask )
                    break;
                idx -= 32;
            }
            prefix = running;
            if ( lane == 0 )
                *reinterpret_cast<volatile ull*>( &status[tile] ) =
                    ( 2ull << 62 ) | ( prefix + agg );
        }
        if ( lane == 0 )
        {
            s_prefix = prefix;
            if ( tile == num_tiles - 1 )
            {
                const ull tot = prefix + agg;
                totals[0] = (long long)( tot & 0x7fffffffull );
                totals[1] = (long long)( ( tot >> 31 ) & 0x7fffffffull );
            }
        }
    }
    __syncthreads();
    const ull excl = s_prefix + s_warp[warp];
    unsigned plo = (unsigned)( excl & 0x7fffffffull );
    unsigned phi = (unsigned)( ( excl >> 31 ) & 0x7fffffffull );
    const unsigned lt = ( 1u << lane ) - 1u;
#pragma unroll
    for ( int e = 0; e < kHaloItems; ++e )
    {
        const unsigned id = (unsigned)( wbase + 32 * e + lane );
        if ( ( blo[e] >> lane ) & 1u )
            steer_lo[plo + (unsigned)__popc( blo[e] & lt )] = id;
        if ( ( bhi[e] >> lane ) & 1u )
            steer_hi[phi + (unsigned)__popc( bhi[e] & lt )] = id;
        plo += (unsigned)__popc( blo[e] );
        phi += (unsigned)__popc( bhi[e] );
    }
}

int make_field_set( const cb_field* fields, int num_fields, FieldSet& fs )
{
    if ( !fields || num_fields < 1 || num_fields > kMaxFields )
        return fail( CB_ERR_INVALID, "comm: need 1..8 fields" );
    // 8-byte members first would be the natural alignment rule; Cabana tuples keep the
    // declared member order, so do we -- and require each member to stay aligned.
    int off = 0;
    for ( int k = 0; k < num_fields; ++k )
    {
        const cb_field& f = fields[k];
        if ( !f.base || f.vlen < 1 || f.num_comp < 1 ||
             ( f.elem_bytes != 4 && f.elem_bytes != 8 ) )
            return fail( CB_ERR_INVALID, "comm: bad field descriptor" );
        if ( off % f.elem_bytes )
            off += f.elem_bytes - off % f.elem_bytes;
        fs.f[k] = make_access( f );
        fs.byte_off[k] = off;
        off += f.num_comp * f.elem_bytes;
    }
    if ( off % 8 )
        off += 8 - off % 8;
    fs.num = num_fields;
    fs.tuple_bytes = off;
    return CB_OK;
}

// Scratch shared by the plan kernels (one caller thread per process, cabana_b200.h).
struct CommScratch
{
    DeviceBuffer counts, flags, slots, scan, dense, idx_a;
    PinnedScalars pinned;
};
CommScratch& scratch()
{
    static thread_local CommScratch s[kMaxDevices];
    return s[current_device_slot()];
}

} // namespace
} // namespace cb

using namespace cb;

extern "C" int cb_slab_halo_select( const cb_positions* x, int64_t num_local,
                                    double lo_thresh, double hi_thresh, int lo_rank,
                                    int hi_rank, int32_t* export_ranks,
                                    uint32_t* export_ids, cb_stream_t stream_ )
{
    if ( !x || !export_ranks || !export_ids || num_local < 0 || num_local > x->n ||
         x->vlen < 1 )
        return fail( CB_ERR_INVALID, "cb_slab_halo_select: bad argument" );
    if ( num_local == 0 )
        return CB_OK;
    k_slab_halo_select<<<launch_grid_for( num_local, kBlock ), kBlock, 0,
                         (cudaStream_t)stream_>>>( make_access( *x ), num_local,
                                                   lo_thresh, hi_thresh, lo_rank,
                                                   hi_rank, export_ranks, export_ids );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_slab_migrate_destinations( const cb_positions* x, int64_t num_local,
                                             const double* bounds_h, int num_ranks,
                                             int32_t* export_ranks, cb_stream_t stream_ )
{
    if ( !x || !bounds_h || !export_ranks || num_local < 0 || num_local > x->n ||
         x->vlen < 1 || num_ranks < 1 || num_ranks >= kMaxBounds )
        return fail( CB_ERR_INVALID, "cb_slab_migrate_destinations: bad argument" );
    if ( num_local == 0 )
        return CB_OK;
    Bounds bd;
    bd.num_ranks = num_ranks;
    for ( int g = 0; g <= num_ranks; ++g )
        bd.b[g] = bounds_h[g];
    k_slab_destinations<<<launch_grid_for( num_local, kBlock ), kBlock, 0,
                          (cudaStream_t)stream_>>>( make_access( *x ), num_local, bd,
                                                    export_ranks );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_comm_count_and_steer( const int32_t* export_ranks, int64_t num_export,
                                        int num_ranks, int64_t* counts_h,
                                        int64_t* offsets_h, uint32_t* steering,
                                        const uint32_t* export_ids, cb_stream_t stream_ )
{
    if ( !counts_h || !offsets_h || num_export < 0 || num_ranks < 1 ||
         num_ranks > kMaxRanks || ( num_export > 0 && ( !export_ranks || !steering ) ) )
        return fail( CB_ERR_INVALID, "cb_comm_count_and_steer: bad argument" );
    cudaStream_t stream = (cudaStream_t)stream_;
    CommScratch& s = scratch();
    CB_TRY( s.pinned.ensure() );
    CB_TRY( s.counts.ensure( sizeof( unsigned long long ) * kMaxRanks ) );
    CB_CUDA( cudaMemsetAsync( s.counts.ptr, 0, sizeof( unsigned long long ) * num_ranks,
                              stream ) );
    // pinned scratch of the plan (64 slots) for small rank counts, else a temporary
    unsigned long long* host_counts =
        reinterpret_cast<unsigned long long*>( s.pinned.ptr );
    unsigned long long* big_counts = nullptr;
    if ( num_ranks > 64 )
    {
        CB_CUDA( cudaMallocHost( (void**)&big_counts,
                                 sizeof( unsigned long long ) * num_ranks ) );
        host_counts = big_counts;
    }
    if ( num_export > 0 )
    {
        k_rank_histogram<<<launch_grid_for( num_export, kBlock * 4 ), kBlock,
                           sizeof( unsigned ) * num_ranks, stream>>>(
            export_ranks, num_export, num_ranks, s.counts.as<unsigned long long>() );
        note_launch();
    }
    cudaError_t e1 = cudaGetLastError();
    if ( e1 == cudaSuccess )
        e1 = cudaMemcpyAsync( host_counts, s.counts.ptr,
                              sizeof( unsigned long long ) * num_ranks,
                              cudaMemcpyDeviceToHost, stream );
    if ( e1 == cudaSuccess )
        e1 = cudaStreamSynchronize( stream );
    if ( e1 != cudaSuccess )
    {
        if ( big_counts )
            cudaFreeHost( big_counts );
        return cuda_fail( e1, "count readback", __FILE__, __LINE__ );
    }
    int64_t run = 0;
    for ( int r = 0; r < num_ranks; ++r )
    {
        counts_h[r] = (int64_t)host_counts[r];
        offsets_h[r] = run;
        run += counts_h[r];
    }
    offsets_h[num_ranks] = run;
    if ( big_counts )
        cudaFreeHost( big_counts );

    if ( num_export > 0 && run > 0 )
    {
        // dense index of the destinations that receive something, in ascending rank order
        std::vector<int> dense( (size_t)num_ranks, -1 );
        int k = 0;
        for ( int r = 0; r < num_ranks; ++r )
            if ( counts_h[r] > 0 )
                dense[(size_t)r] = k++;
        const long long tiles0 = ( num_export + kSteerTile - 1 ) / kSteerTile;
        CB_TRY( s.dense.ensure( sizeof( int ) * (size_t)kMaxRanks ) );
        CB_TRY( s.flags.ensure( sizeof( int ) * (size_t)( kSteerDigits * tiles0 ), 1.1 ) );
        CB_TRY( s.slots.ensure( sizeof( int ) * (size_t)( kSteerDigits * tiles0 + 1 ), 1.1 ) );
        CB_CUDA( cudaMemcpyAsync( s.dense.ptr, dense.data(), sizeof( int ) * (size_t)num_ranks,
                                  cudaMemcpyHostToDevice, stream ) );
        CB_CUDA( cudaStreamSynchronize( stream ) ); // (dense is a stack-lifetime host buffer)
        const int* dense_d = s.dense.as<int>();
        int* hist = s.flags.as<int>();
        int* base = s.slots.as<int>();
        const bool two = k > kSteerDigits;
        if ( two )
            CB_TRY( s.idx_a.ensure( sizeof( unsigned ) * (size_t)run, 1.1 ) );
        // pass over the low 5 bits (the only one for <= 32 destinations)
        k_steer_digit_hist<true><<<(unsigned)tiles0, kSteerTile, 0, stream>>>(
            export_ranks, nullptr, num_export, dense_d, 0, hist, tiles0 );
        CB_CHECK_LAUNCH();
        CB_TRY( exclusive_scan_i32( hist, base, kSteerDigits * tiles0, false, nullptr, s.scan,
                                    stream ) );
        if ( !two )
            k_steer_digit_scatter<true, true><<<(unsigned)tiles0, kSteerTile, 0, stream>>>(
                export_ranks, nullptr, num_export, dense_d, 0, base, tiles0, export_ids,
                steering );
        else
            k_steer_digit_scatter<true, false><<<(unsigned)tiles0, kSteerTile, 0, stream>>>(
                export_ranks, nullptr, num_export, dense_d, 0, base, tiles0, nullptr,
                s.idx_a.as<unsigned>() );
        CB_CHECK_LAUNCH();
        if ( two )
        {
            // pass over the high 5 bits of what the first pass kept (stable: LSD radix)
            const long long tiles1 = ( run + kSteerTile - 1 ) / kSteerTile;
            k_steer_digit_hist<false><<<(unsigned)tiles1, kSteerTile, 0, stream>>>(
                export_ranks, s.idx_a.as<unsigned>(), run, dense_d, 5, hist, tiles1 );
            CB_CHECK_LAUNCH();
            CB_TRY( exclusive_scan_i32( hist, base, kSteerDigits * tiles1, false, nullptr,
                                        s.scan, stream ) );
            k_steer_digit_scatter<false, true><<<(unsigned)tiles1, kSteerTile, 0, stream>>>(
                export_ranks, s.idx_a.as<unsigned>(), run, dense_d, 5, base, tiles1, export_ids,
                steering );
            CB_CHECK_LAUNCH();
        }
    }
    return CB_OK;
}

extern "C" int64_t cb_comm_tuple_bytes( const cb_field* fields, int num_fields )
{
    FieldSet fs;
    if ( make_field_set( fields, num_fields, fs ) != CB_OK )
        return -1;
    return fs.tuple_bytes;
}

extern "C" int cb_comm_pack( const cb_field* fields, int num_fields,
                             const uint32_t* steering, int64_t count, void* send_buffer,
                             cb_stream_t stream_ )
{
    FieldSet fs;
    CB_TRY( make_field_set( fields, num_fields, fs ) );
    if ( count < 0 || ( count > 0 && ( !steering || !send_buffer ) ) )
        return fail( CB_ERR_INVALID, "cb_comm_pack: bad argument" );
    if ( count == 0 )
        return CB_OK;
    k_pack_unpack<0><<<launch_grid_for( count, kBlock ), kBlock, 0,
                       (cudaStream_t)stream_>>>( fs, steering, 0, count,
                                                 (char*)send_buffer );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_comm_unpack( const cb_field* fields, int num_fields, int64_t dst_begin,
                               int64_t count, const void* recv_buffer,
                               cb_stream_t stream_ )
{
    FieldSet fs;
    CB_TRY( make_field_set( fields, num_fields, fs ) );
    if ( count < 0 || dst_begin < 0 || ( count > 0 && !recv_buffer ) )
        return fail( CB_ERR_INVALID, "cb_comm_unpack: bad argument" );
    for ( int k = 0; k < num_fields; ++k )
        if ( fields[k].n < dst_begin + count )
            return fail( CB_ERR_INVALID, "cb_comm_unpack: field too small for ghosts" );
    if ( count == 0 )
        return CB_OK;
    k_pack_unpack<1><<<launch_grid_for( count, kBlock ), kBlock, 0,
                       (cudaStream_t)stream_>>>( fs, nullptr, dst_begin, count,
                                                 (char*)recv_buffer );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_comm_scatter_add_typed( const cb_field* field, const uint32_t* steering,
                                          int64_t count, const void* recv_buffer, int dtype,
                                          cb_stream_t stream_ )
{
    static const int kBytes[4] = { 8, 4, 4, 8 };
    if ( !field || dtype < CB_DTYPE_F64 || dtype > CB_DTYPE_I64 || field->vlen < 1 ||
         field->num_comp < 1 || count < 0 || ( count > 0 && ( !steering || !recv_buffer ) ) )
        return fail( CB_ERR_INVALID, "cb_comm_scatter_add: bad argument" );
    if ( field->elem_bytes != kBytes[dtype] )
        return fail( CB_ERR_INVALID, "cb_comm_scatter_add: element size does not match dtype" );
    if ( count == 0 )
        return CB_OK;
    const int grid = launch_grid_for( count, kBlock );
    cudaStream_t stream = (cudaStream_t)stream_;
    const FieldAccess f = make_access( *field );
    const int te = (int)( cb_comm_tuple_bytes( field, 1 ) / field->elem_bytes );
    switch ( dtype )
    {
    case CB_DTYPE_F64:
        k_scatter_add<double><<<grid, kBlock, 0, stream>>>( f, steering, count,
                                                            (const double*)recv_buffer, te );
        break;
    case CB_DTYPE_F32:
        k_scatter_add<float><<<grid, kBlock, 0, stream>>>( f, steering, count,
                                                           (const float*)recv_buffer, te );
        break;
    case CB_DTYPE_I32:
        k_scatter_add<int><<<grid, kBlock, 0, stream>>>( f, steering, count,
                                                         (const int*)recv_buffer, te );
        break;
    default: // two's complement: the unsigned 64-bit atomic adds signed values correctly
        k_scatter_add<unsigned long long><<<grid, kBlock, 0, stream>>>(
            f, steering, count, (const unsigned long long*)recv_buffer, te );
        break;
    }
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_comm_scatter_add( const cb_field* field, const uint32_t* steering,
                                    int64_t count, const void* recv_buffer,
                                    cb_stream_t stream_ )
{
    if ( field && field->elem_bytes != 8 )
        return fail( CB_ERR_INVALID, "cb_comm_scatter_add: doubles only (use "
                                     "cb_comm_scatter_add_typed for other value types)" );
    return cb_comm_scatter_add_typed( field, steering, count, recv_buffer, CB_DTYPE_F64,
                                      stream_ );
}

extern "C" int cb_comm_pack_range( const cb_field* fields, int num_fields, int64_t src_begin,
                                   int64_t count, void* send_buffer, cb_stream_t stream_ )
{
    FieldSet fs;
    CB_TRY( make_field_set( fields, num_fields, fs ) );
    if ( count < 0 || src_begin < 0 || ( count > 0 && !send_buffer ) )
        return fail( CB_ERR_INVALID, "cb_comm_pack_range: bad argument" );
    for ( int k = 0; k < num_fields; ++k )
        if ( fields[k].n < src_begin + count )
            return fail( CB_ERR_INVALID, "cb_comm_pack_range: range exceeds the field" );
    if ( count == 0 )
        return CB_OK;
    k_pack_unpack<2><<<launch_grid_for( count, kBlock ), kBlock, 0,
                       (cudaStream_t)stream_>>>( fs, nullptr, src_begin, count,
                                                 (char*)send_buffer );
    CB_CHECK_LAUNCH();
    return CB_OK;
}

extern "C" int cb_slab_halo_plan( const cb_positions* x, int64_t num_local, double lo_thresh,
                                  double hi_thresh, int has_lo, int has_hi, uint32_t* steer_lo,
                                  uint32_t* steer_hi, int64_t* counts_h, cb_stream_t stream_ )
{
    if ( !x || !steer_lo || !steer_hi || !counts_h || num_local < 0 || num_local > x->n ||
         x->vlen < 1 || num_local >= 2147483647ll )
        return fail( CB_ERR_INVALID, "cb_slab_halo_plan: bad argument" );
    counts_h[0] = counts_h[1] = 0;
    if ( num_local == 0 )
        return CB_OK;
    cudaStream_t stream = (cudaStream_t)stream_;
    CommScratch& s = scratch();
    CB_TRY( s.pinned.ensure() );
    const long long num_tiles = ( num_local + kHaloTile - 1 ) / kHaloTile;
    const size_t bytes = 32 + (size_t)num_tiles * sizeof( unsigned long long );
    CB_TRY( s.scan.ensure( bytes, 1.25 ) );
    CB_CUDA( cudaMemsetAsync( s.scan.ptr, 0, bytes, stream ) );
    unsigned* counter = s.scan.as<unsigned>();
    long long* totals = reinterpret_cast<long long*>( s.scan.as<char>() + 16 );
    unsigned long long* status = reinterpret_cast<unsigned long long*>( s.scan.as<char>() + 32 );
    k_halo_compact<<<(unsigned)num_tiles, kHaloThreads, 0, stream>>>(
        make_access( *x ), num_local, lo_thresh, hi_thresh, has_lo, has_hi, steer_lo, steer_hi,
        status, counter, (unsigned)num_tiles, totals );
    CB_CHECK_LAUNCH();
    CB_CUDA( cudaMemcpyAsync( s.pinned.ptr, totals, 2 * sizeof( long long ),
                              cudaMemcpyDeviceToHost, stream ) );
    CB_CUDA( cudaStreamSynchronize( stream ) );
    counts_h[0] = s.pinned.ptr[0];
    counts_h[1] = s.pinned.ptr[1];
    return CB_OK;
}

// =====================================================================================
// Peer-memory slab halo (CUDA IPC over NVLink / NVSwitch)
// =====================================================================================
namespace cb
{
namespace
{
constexpr size_t kWinHeader = 128; // per buffer: u64 sequence, i64 count

struct WinLayout
{
    long long capacity;
    long long tuple_bytes;
    size_t buffer_bytes() const { return kWinHeader + (size_t)capacity * (size_t)tuple_bytes; }
};

// pack the export list of one side into a (possibly remote) window buffer
// Both faces in ONE launch (blockIdx.y = face), and the publication folded in: every CTA fences
// its tuples system-wide and takes a ticket; the face's last CTA writes the count, then the
// sequence number, into the neighbour's window header (what k_halo_signal does on its own).
struct PushFace
{
    const unsigned* steering;
    char* buf; // peer window buffer of this sequence parity (header + data), nullptr: no neighbour
};
__global__ void __launch_bounds__( kBlock )
    k_halo_push_signal( FieldSet fs, PushFace lo, PushFace hi, const long long* __restrict__ totals,
                        long long capacity, unsigned* done /* [2], zeroed */,
                        unsigned long long seq )
{
    const int side = (int)blockIdx.y;
    const PushFace face = side ? hi : lo;
    if ( !face.buf )
        return;
    const long long count = min( totals[side], capacity );
    char* data = face.buf + kWinHeader;
    for ( long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < count;
          j += (long long)gridDim.x * kBlock )
    {
        const long long elem = (long long)face.steering[j];
        char* tuple = data + j * fs.tuple_bytes;
        for ( int k = 0; k < fs.num; ++k )
        {
            const FieldAccess& f = fs.f[k];
            const long long off = f.offset( elem );
            if ( f.elem_bytes == 8 )
            {
                const unsigned long long* fb = reinterpret_cast<const unsigned long long*>( f.base );
                unsigned long long* tb = reinterpret_cast<unsigned long long*>( tuple + fs.byte_off[k] );
                for ( int c = 0; c < f.num_comp; ++c )
                    tb[c] = fb[off + f.comp_stride * c];
            }
            else
            {
                const unsigned* fb = reinterpret_cast<const unsigned*>( f.base );
                unsigned* tb = reinterpret_cast<unsigned*>( tuple + fs.byte_off[k] );
                for ( int c = 0; c < f.num_comp; ++c )
                    tb[c] = fb[off + f.comp_stride * c];
            }
        }
    }
    __threadfence_system();
    __syncthreads();
    if ( threadIdx.x == 0 )
    {
        const unsigned ticket = atomicAdd( done + side, 1u );
        if ( ticket == gridDim.x - 1 )
        {
            __threadfence_system();
            *reinterpret_cast<volatile long long*>( face.buf + 8 ) = totals[side];
            __threadfence_system();
            *reinterpret_cast<volatile unsigned long long*>( face.buf ) = seq;
            __threadfence_system();
        }
    }
}

// receiver: bounded wait on the local header(s)
__global__ void k_halo_wait( const char* hdr_lo, const char* hdr_hi, unsigned long long seq,
                             long long* out /* [3]: count_lo, count_hi, timed_out */ )
{
    const long long start = clock64();
    const long long limit = 20000000000ll; // ~10 s at 2 GHz: never hang the GPU
    bool ok = true;
    for ( int side = 0; side < 2; ++side )
    {
        const char* h = side ? hdr_hi : hdr_lo;
        long long count = 0;
        if ( h )
        {
            while ( *reinterpret_cast<const volatile unsigned long long*>( h ) != seq )
            {
                if ( clock64() - start > limit )
                {
                    ok = false;
                    break;
                }
                __nanosleep( 200 );
            }
            __threadfence_system();
            count = *reinterpret_cast<const volatile long long*>( h + 8 );
        }
        out[side] = count;
    }
    out[2] = ok ? 0 : 1;
}

// Receiver, no host in the loop: every CTA waits (bounded) for the neighbours' publications on
// its OWN window headers, then the grid unpacks both ghost layers behind the owned particles
// (the lower neighbour's first: impl/Cabana_Halo_Mpi.hpp:108-121) and leaves the counts, the
// flags and the new particle count in device memory for the build that follows on the stream.
// blockIdx.y = face.  out[0..1] = counts as published, out[2] = flags (1 timed out, 2 a layer
// larger than the window, 4 the slices cannot hold the ghosts), out[3] = num_local + ghosts.
__global__ void __launch_bounds__( kBlock )
    k_halo_wait_unpack( FieldSet fs, const char* hdr_lo, const char* hdr_hi,
                        unsigned long long seq, long long num_local, long long capacity,
                        long long field_n, long long* out )
{
    __shared__ long long s_cnt[2];
    __shared__ int s_flags;
    if ( threadIdx.x == 0 )
    {
        const long long start = clock64();
        const long long limit = 20000000000ll; // ~10 s at 2 GHz: never hang the GPU
        int flags = 0;
        for ( int side = 0; side < 2; ++side )
        {
            const char* h = side ? hdr_hi : hdr_lo;
            long long count = 0;
            if ( h )
            {
                while ( *reinterpret_cast<const volatile unsigned long long*>( h ) != seq )
                {
                    if ( clock64() - start > limit )
                    {
                        flags |= 1;
                        break;
                    }
                    __nanosleep( 100 );
                }
                __threadfence_system();
                if ( !( flags & 1 ) )
                    count = *reinterpret_cast<const volatile long long*>( h + 8 );
            }
            s_cnt[side] = count;
        }
        s_flags = flags;
    }
    __syncthreads();
    int flags = s_flags;
    long long c_lo = s_cnt[0], c_hi = s_cnt[1];
    if ( c_lo > capacity || c_hi > capacity )
        flags |= 2;
    c_lo = min( max( c_lo, 0ll ), capacity );
    c_hi = min( max( c_hi, 0ll ), capacity );
    const long long room = max( field_n - num_local, 0ll );
    if ( c_lo + c_hi > room )
    {
        flags |= 4;
        c_lo = min( c_lo, room );
        c_hi = min( c_hi, room - c_lo );
    }
    if ( blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0 )
    {
        out[0] = s_cnt[0];
        out[1] = s_cnt[1];
        out[2] = flags;
        out[3] = num_local + c_lo + c_hi;
    }
    const int side = (int)blockIdx.y;
    const char* h = side ? hdr_hi : hdr_lo;
    if ( !h )
        return;
    const char* data = h + kWinHeader;
    const long long dst_begin = num_local + ( side ? c_lo : 0 );
    const long long count = side ? c_hi : c_lo;
    for ( long long j = (long long)blockIdx.x * kBlock + threadIdx.x; j < count;
          j += (long long)gridDim.x * kBlock )
    {
        const long long elem = dst_begin + j;
        const char* tuple = data + j * fs.tuple_bytes;
        for ( int k = 0; k < fs.num; ++k )
        {
            const FieldAccess& f = fs.f[k];
            const long long off = f.offset( elem );
            // (L2 loads: the tuples were written by the neighbour over NVLink)
            if ( f.elem_bytes == 8 )
            {
                unsigned long long* fb = reinterpret_cast<unsigned long long*>( f.base );
                const unsigned long long* tb =
                    reinterpret_cast<const unsigned long long*>( tuple + fs.byte_off[k] );
                for ( int c = 0; c < f.num_comp; ++c )
                    fb[off + f.comp_stride * c] = __ldcg( tb + c );
            }
            else
            {
                unsigned* fb = reinterpret_cast<unsigned*>( f.base );
                const unsigned* tb = reinterpret_cast<const unsigned*>( tuple + fs.byte_off[k] );
                for ( int c = 0; c < f.num_comp; ++c )
                    fb[off + f.comp_stride * c] = __ldcg( tb + c );
            }
        }
    }
}
} // namespace
} // namespace cb

struct cb_p2p_window
{
    char* base = nullptr; // two buffers: [header|data][header|data]
    cb::WinLayout lay{};
    cb::DeviceBuffer scratch; // wait results
};

extern "C" int cb_p2p_window_create( cb_p2p_window** out, int64_t capacity_tuples,
                                     int64_t tuple_bytes )
{
    if ( !out || capacity_tuples < 1 || tuple_bytes < 1 )
        return fail( CB_ERR_INVALID, "cb_p2p_window_create: bad argument" );
    cb_p2p_window* w = new ( std::nothrow ) cb_p2p_window();
    if ( !w )
        return fail( CB_ERR_NOMEM, "cb_p2p_window_create" );
    w->lay.capacity = capacity_tuples;
    w->lay.tuple_bytes = tuple_bytes;
    const size_t bytes = 2 * w->lay.buffer_bytes();
    cudaError_t e = cudaMalloc( (void**)&w->base, bytes ); // plain cudaMalloc: IPC-exportable
    if ( e == cudaSuccess )
        e = cudaMemset( w->base, 0, bytes );
    // (cudaMemset returns before the legacy stream has run it, and a caller on a non-blocking
    // stream is not ordered behind that stream: the headers must be zero before a handle leaves)
    if ( e == cudaSuccess )
        e = cudaDeviceSynchronize();
    if ( e != cudaSuccess )
    {
        delete w;
        return cuda_fail( e, "cb_p2p_window_create", __FILE__, __LINE__ );
    }
    *out = w;
    return CB_OK;
}

extern "C" int cb_p2p_window_destroy( cb_p2p_window* w )
{
    if ( w )
    {
        if ( w->base )
            cudaFree( w->base );
        delete w;
    }
    return CB_OK;
}

extern "C" int cb_p2p_window_get_handle( const cb_p2p_window* w, void* handle_h )
{
    if ( !w || !handle_h )
        return fail( CB_ERR_INVALID, "cb_p2p_window_get_handle: null argument" );
    static_assert( sizeof( cudaIpcMemHandle_t ) == CB_IPC_HANDLE_BYTES, "" );
    cudaIpcMemHandle_t h;
    CB_CUDA( cudaIpcGetMemHandle( &h, w->base ) );
    memcpy( handle_h, &h, sizeof( h ) );
    return CB_OK;
}

extern "C" int cb_p2p_window_local_base( const cb_p2p_window* w, void** base )
{
    if ( !w || !base )
        return fail( CB_ERR_INVALID, "cb_p2p_window_local_base: null argument" );
    *base = w->base;
    return CB_OK;
}

extern "C" int cb_p2p_window_open( const void* handle_h, void** peer_base )
{
    if ( !handle_h || !peer_base )
        return fail( CB_ERR_INVALID, "cb_p2p_window_open: null argument" );
    cudaIpcMemHandle_t h;
    memcpy( &h, handle_h, sizeof( h ) );
    CB_CUDA( cudaIpcOpenMemHandle( peer_base, h, cudaIpcMemLazyEnablePeerAccess ) );
    return CB_OK;
}

extern "C" int cb_p2p_window_close( void* peer_base )
{
    if ( peer_base )
        CB_CUDA( cudaIpcCloseMemHandle( peer_base ) );
    return CB_OK;
}

extern "C" int cb_slab_halo_push( const cb_positions* x, const cb_field* fields,
                                  int num_fields, int64_t num_local, double lo_thresh,
                                  double hi_thresh, void* peer_lo, void* peer_hi,
                                  int64_t capacity_tuples, uint64_t sequence,
                                  uint32_t* steer_scratch, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::gather" );
    FieldSet fs;
    CB_TRY( make_field_set( fields, num_fields, fs ) );
    if ( !x || !steer_scratch || num_local < 0 || num_local > x->n || x->vlen < 1 ||
         num_local >= 2147483647ll || capacity_tuples < 1 )
        return fail( CB_ERR_INVALID, "cb_slab_halo_push: bad argument" );
    cudaStream_t stream = (cudaStream_t)stream_;
    CommScratch& s = scratch();
    const long long nl = num_local > 0 ? num_local : 1;
    const long long num_tiles = ( nl + kHaloTile - 1 ) / kHaloTile;
    const size_t bytes = 32 + (size_t)num_tiles * sizeof( unsigned long long );
    CB_TRY( s.scan.ensure( bytes, 1.25 ) );
    CB_CUDA( cudaMemsetAsync( s.scan.ptr, 0, bytes, stream ) );
    unsigned* counter = s.scan.as<unsigned>();
    long long* totals = reinterpret_cast<long long*>( s.scan.as<char>() + 16 );
    unsigned long long* status = reinterpret_cast<unsigned long long*>( s.scan.as<char>() + 32 );
    uint32_t* steer_lo = steer_scratch;
    uint32_t* steer_hi = steer_scratch + nl;
    if ( num_local > 0 )
    {
        k_halo_compact<<<(unsigned)num_tiles, kHaloThreads, 0, stream>>>(
            make_access( *x ), num_local, lo_thresh, hi_thresh, peer_lo ? 1 : 0,
            peer_hi ? 1 : 0, steer_lo, steer_hi, status, counter, (unsigned)num_tiles, totals );
        CB_CHECK_LAUNCH();
    }
    WinLayout lay{ capacity_tuples, fs.tuple_bytes };
    const size_t parity_off = ( sequence & 1ull ) ? lay.buffer_bytes() : 0;
    // my LOWER neighbour receives my low-face particles in ITS "from_hi" window and vice
    // versa; the caller passes the matching peer bases, so the face only selects the list.
    // Ghost layers are thin: a grid sized for the capacity, threads past the count exit.
    if ( peer_lo || peer_hi )
    {
        const PushFace flo{ steer_lo, peer_lo ? (char*)peer_lo + parity_off : nullptr };
        const PushFace fhi{ steer_hi, peer_hi ? (char*)peer_hi + parity_off : nullptr };
        // (one CTA per SM and face, grid-stride: every CTA pays one system-wide fence and one
        // ticket however many tuples it moved -- the layer's size is only known on the device)
        long long push_ctas = launch_grid_for( capacity_tuples, kBlock );
        if ( push_ctas > kNumSMs )
            push_ctas = kNumSMs;
        const dim3 grid( (unsigned)push_ctas, 2u );
        k_halo_push_signal<<<grid, kBlock, 0, stream>>>(
            fs, flo, fhi, totals, capacity_tuples, s.scan.as<unsigned>() + 2,
            (unsigned long long)sequence );
        CB_CHECK_LAUNCH();
    }
    return CB_OK;
}

extern "C" int cb_slab_halo_wait( cb_p2p_window* from_lo, cb_p2p_window* from_hi,
                                  uint64_t sequence, int64_t* counts_h, const void** data_lo,
                                  const void** data_hi, cb_stream_t stream_ )
{
    ScopedRegion region( "Cabana::gather" );
    if ( !counts_h || !data_lo || !data_hi )
        return fail( CB_ERR_INVALID, "cb_slab_halo_wait: null argument" );
    cudaStream_t stream = (cudaStream_t)stream_;
    counts_h[0] = counts_h[1] = 0;
    *data_lo = *data_hi = nullptr;
    if ( !from_lo && !from_hi )
        return CB_OK;
    CommScratch& s = scratch();
    CB_TRY( s.pinned.ensure() );
    cb_p2p_window* any = from_lo ? from_lo : from_hi;
    CB_TRY( any->scratch.ensure( 64 ) );
    const char* hl = nullptr;
    const char* hh = nullptr;
    if ( from_lo )
    {
        hl = from_lo->base + ( ( sequence & 1ull ) ? from_lo->lay.buffer_bytes() : 0 );
        *data_lo = hl + kWinHeader;
    }
    if ( from_hi )
    {
        hh = from_hi->base + ( ( sequence & 1ull ) ? from_hi->lay.buffer_bytes() : 0 );
        *data_hi = hh + kWinHeader;
    }
    k_halo_wait<<<1, 1, 0, stream>>>( hl, hh, (unsigned long long)sequence,
                                      any->scratch.as<long long>() );
    CB_CHECK_LAUNCH();
    CB_CUDA( cudaMemcpyAsync( s.pinned.ptr + 8, any->scratch.ptr, 3 * sizeof( long long ),
                              cudaMemcpyDeviceToHost, stream ) );
    CB_CUDA( cudaStreamSynchronize( stream ) );
    if ( s.pinned.ptr[10] != 0 )
        return fail( CB_ERR_CUDA, "cb_slab_halo_wait: timed out waiting for a neighbour's push" );
    counts_h[0] = s.pinned.ptr[8];
    counts_h[1] = s.pinned.ptr[9];
    if ( ( from_lo && counts_h[0] > from_lo->lay.capacity ) ||
         ( from_hi && counts_h[1] > from_hi->lay.capacity ) )
        return fail( CB_ERR_NOMEM, "cb_slab_halo_wait: ghost layer larger than the window" );
    return CB_OK;
}

// One host entry for the sharded step: ghost selection + plan + push to the neighbours' windows
// -> wait for theirs -> unpack behind the owned particles -> owner-local Verlet build with
// begin = 0, end = num_local (impl/Cabana_Halo_Mpi.hpp:41-125 + Cabana_VerletList.hpp:1351-1392).
// Two host synchronisations in all (the ghost counts; the list size).
extern "C" int cb_slab_step( cb_verlet* list, const cb_positions* x_all, const cb_field* fields,
                             int num_fields, int64_t num_local, double lo_thresh,
                             double hi_thresh, void* peer_lo, void* peer_hi,
                             cb_p2p_window* from_lo, cb_p2p_window* from_hi,
                             int64_t capacity_tuples, uint64_t sequence,
                             uint32_t* steer_scratch, double radius, double cell_size_ratio,
                             const double* grid_min, const double* grid_max, int64_t max_neigh,
                             int algorithm, int layout, int build_op, int64_t* counts_h,
                             cb_stream_t stream )
{
    if ( !list || !x_all || !fields || !counts_h || num_local < 0 || num_local > x_all->n )
        return fail( CB_ERR_INVALID, "cb_slab_step: bad argument" );
    cb_positions x_own = *x_all;
    x_own.n = num_local;
    CB_TRY( cb_slab_halo_push( &x_own, fields, num_fields, num_local, lo_thresh, hi_thresh,
                               peer_lo, peer_hi, capacity_tuples, sequence, steer_scratch,
                               stream ) );
    counts_h[0] = counts_h[1] = 0;
    const char* sync_env = getenv( "CB_SLAB_SYNC" );
    const bool device_counts = ( from_lo || from_hi ) && cb::verlet_devcount_supported() &&
                               !( sync_env && sync_env[0] == '1' );
    if ( device_counts )
    {
        // Ghost counts never visit the host before the build: wait + unpack is one kernel that
        // leaves the particle count on the device, the build reads it there, and the counts come
        // back with the build's own (single, overlapped) read-back.
        ScopedRegion region( "Cabana::gather" );
        FieldSet fs;
        CB_TRY( make_field_set( fields, num_fields, fs ) );
        long long field_n = x_all->n;
        for ( int k = 0; k < num_fields; ++k )
            field_n = fields[k].n < field_n ? fields[k].n : field_n;
        cudaStream_t st = (cudaStream_t)stream;
        CommScratch& s = scratch();
        CB_TRY( s.pinned.ensure() );
        cb_p2p_window* any = from_lo ? from_lo : from_hi;
        CB_TRY( any->scratch.ensure( 64 ) );
        const size_t poff_lo =
            from_lo ? ( ( sequence & 1ull ) ? from_lo->lay.buffer_bytes() : 0 ) : 0;
        const size_t poff_hi =
            from_hi ? ( ( sequence & 1ull ) ? from_hi->lay.buffer_bytes() : 0 ) : 0;
        const char* hl = from_lo ? from_lo->base + poff_lo : nullptr;
        const char* hh = from_hi ? from_hi->base + poff_hi : nullptr;
        const long long cap = from_lo ? from_lo->lay.capacity : from_hi->lay.capacity;
        if ( ( from_lo && from_lo->lay.tuple_bytes != fs.tuple_bytes ) ||
             ( from_hi && ( from_hi->lay.tuple_bytes != fs.tuple_bytes ||
                            from_hi->lay.capacity != cap ) ) )
            return fail( CB_ERR_INVALID, "cb_slab_step: windows do not match the fields" );
        long long* out = any->scratch.as<long long>();
        const dim3 grid( (unsigned)launch_grid_for( cap, kBlock ), 2u );
        k_halo_wait_unpack<<<grid, kBlock, 0, st>>>( fs, hl, hh, (unsigned long long)sequence,
                                                     num_local, cap, field_n, out );
        CB_CHECK_LAUNCH();
        CB_CUDA( cudaMemcpyAsync( s.pinned.ptr + 8, out, 4 * sizeof( long long ),
                                  cudaMemcpyDeviceToHost, st ) );
        cb_positions x_tot = *x_all;
        const long long bound = num_local + ( from_lo ? cap : 0 ) + ( from_hi ? cap : 0 );
        x_tot.n = bound < field_n ? bound : field_n;
        const int rc = cb::verlet_build_devcount( list, &x_tot, 0, num_local, radius,
                                                  cell_size_ratio, grid_min, grid_max, max_neigh,
                                                  algorithm, layout, build_op, stream, out + 3 );
        if ( rc != CB_OK )
            return rc;
        // (the build waited for an event recorded after the copy above)
        const long long flags = s.pinned.ptr[10];
        if ( flags & 1 )
            return fail( CB_ERR_CUDA, "cb_slab_step: timed out waiting for a neighbour's push" );
        if ( flags & 2 )
            return fail( CB_ERR_NOMEM, "cb_slab_step: ghost layer larger than the window" );
        if ( flags & 4 )
            return fail( CB_ERR_NOMEM, "cb_slab_step: position slice too small for the ghosts" );
        counts_h[0] = s.pinned.ptr[8];
        counts_h[1] = s.pinned.ptr[9];
        return CB_OK;
    }
    const void* data_lo = nullptr;
    const void* data_hi = nullptr;
    CB_TRY( cb_slab_halo_wait( from_lo, from_hi, sequence, counts_h, &data_lo, &data_hi,
                               stream ) );
    const int64_t n_tot = num_local + counts_h[0] + counts_h[1];
    if ( n_tot > x_all->n )
        return fail( CB_ERR_NOMEM, "cb_slab_step: position slice too small for the ghosts" );
    if ( counts_h[0] > 0 )
        CB_TRY( cb_comm_unpack( fields, num_fields, num_local, counts_h[0], data_lo, stream ) );
    if ( counts_h[1] > 0 )
        CB_TRY( cb_comm_unpack( fields, num_fields, num_local + counts_h[0], counts_h[1],
                                data_hi, stream ) );
    cb_positions x_tot = *x_all;
    x_tot.n = n_tot;
    return cb_verlet_build( list, &x_tot, 0, num_local, radius, cell_size_ratio, grid_min,
                            grid_max, max_neigh, algorithm, layout, build_op, stream );
}
