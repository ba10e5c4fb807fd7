// Internal host-side helpers shared by the .cu translation units.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

#include <nvtx3/nvToolsExt.h>

#include "../../include/cabana_b200.h"

namespace cb
{

// Thread-local last error text (cb_last_error_string()).
void set_error( const std::string& msg );
int fail( cb_status code, const char* what );
int cuda_fail( cudaError_t err, const char* what, const char* file, int line );

#define CB_CUDA( expr )                                                        \
    do                                                                         \
    {                                                                          \
        cudaError_t cb_err__ = ( expr );                                       \
        if ( cb_err__ != cudaSuccess )                                         \
            return ::cb::cuda_fail( cb_err__, #expr, __FILE__, __LINE__ );     \
    } while ( 0 )

#define CB_TRY( expr )                                                         \
    do                                                                         \
    {                                                                          \
        int cb_rc__ = ( expr );                                                \
        if ( cb_rc__ != CB_OK )                                                \
            return cb_rc__;                                                    \
    } while ( 0 )

// Tracing range with the reference's region names (Kokkos::Profiling::ScopedRegion in
// Cabana_VerletList.hpp:1377, Cabana_LinkedCellList.hpp:655, Cabana_Parallel.hpp:259,
// Cabana_Sort.hpp:557, impl/Cabana_Halo_Mpi.hpp:48): NVTX, visible to nsys / ncu --nvtx.
struct ScopedRegion
{
    explicit ScopedRegion( const char* name ) { nvtxRangePushA( name ); }
    ~ScopedRegion() { nvtxRangePop(); }
    ScopedRegion( const ScopedRegion& ) = delete;
    ScopedRegion& operator=( const ScopedRegion& ) = delete;
};

// Scratch that outlives a call is kept PER DEVICE (function attributes and allocations belong
// to the device that was current when they were made; a process may cb_set_device() between
// calls).
constexpr int kMaxDevices = 64;
inline int current_device_slot()
{
    int d = 0;
    if ( cudaGetDevice( &d ) != cudaSuccess || d < 0 )
        d = 0;
    return d < kMaxDevices ? d : kMaxDevices - 1;
}

void note_launch();
#define CB_CHECK_LAUNCH()                                                      \
    do                                                                         \
    {                                                                          \
        ::cb::note_launch();                                                   \
        CB_CUDA( cudaGetLastError() );                                         \
    } while ( 0 )

// Grow-only device buffer: rebuilds reuse memory (no cudaMalloc in steady state).
struct DeviceBuffer
{
    void* ptr = nullptr;
    size_t capacity = 0;

    DeviceBuffer() = default;
    DeviceBuffer( const DeviceBuffer& ) = delete;
    DeviceBuffer& operator=( const DeviceBuffer& ) = delete;
    ~DeviceBuffer() { release(); }

    void swap( DeviceBuffer& o )
    {
        void* p = ptr;
        ptr = o.ptr;
        o.ptr = p;
        const size_t c = capacity;
        capacity = o.capacity;
        o.capacity = c;
    }

    // Ensure at least `bytes`; contents are NOT preserved on growth.
    int ensure( size_t bytes, double growth = 1.0 )
    {
        if ( bytes <= capacity )
            return CB_OK;
        release();
        size_t want = (size_t)( (double)bytes * growth );
        if ( want < bytes )
            want = bytes;
        if ( want < 256 )
            want = 256;
        cudaError_t err = cudaMalloc( &ptr, want );
        if ( err != cudaSuccess && want > bytes )
        {
            (void)cudaGetLastError();
            want = bytes;
            err = cudaMalloc( &ptr, want );
        }
        if ( err != cudaSuccess )
        {
            ptr = nullptr;
            capacity = 0;
            return cuda_fail( err, "cudaMalloc", __FILE__, __LINE__ );
        }
        capacity = want;
        return CB_OK;
    }
    void release()
    {
        if ( ptr )
            cudaFree( ptr );
        ptr = nullptr;
        capacity = 0;
    }
    template <class T>
    T* as() const
    {
        return reinterpret_cast<T*>( ptr );
    }
};

// Pinned host scratch for scalar read-backs.
struct PinnedScalars
{
    long long* ptr = nullptr;
    int ensure()
    {
        if ( ptr )
            return CB_OK;
        cudaError_t err = cudaMallocHost( (void**)&ptr, 64 * sizeof( long long ) );
        if ( err != cudaSuccess )
            return cuda_fail( err, "cudaMallocHost", __FILE__, __LINE__ );
        return CB_OK;
    }
    ~PinnedScalars()
    {
        if ( ptr )
            cudaFreeHost( ptr );
    }
};

// ---- cb_scan.cu ------------------------------------------------------------------
// Single-pass decoupled look-back exclusive scan of n int32 values.
//   out[i] = sum_{j<i} in[i] (int32; callers check the 64-bit total for overflow),
//   out[n] = total when write_total_at_n, *total_dev (long long, device) = total.
// `scratch` is grown as needed.  in == out is allowed.
int exclusive_scan_i32( const int* in, int* out, long long n, bool write_total_at_n,
                        long long* total_dev, DeviceBuffer& scratch,
                        cudaStream_t stream );
// stats_dev[0] = max(in), stats_dev[1] = sum(in) (long long, device).
int exclusive_scan_stats_i32( const int* in, int* out, long long n, long long* stats_dev,
                              DeviceBuffer& scratch, cudaStream_t stream );
int max_and_sum_i32( const int* in, long long n, long long* stats_dev,
                     cudaStream_t stream );

// ---- cb_lcl.cu -------------------------------------------------------------------
struct Grid;
struct PosAccess;

// Bin particles [begin,end) on `grid` (each cell split into refine^3 sub-cells when
// refine > 1; cell ids are then cardinal indices of the refined grid):
//   counts[ncell], offsets[ncell+1] (uint32), permute[end-begin] = absolute ids,
//   cell_of[end-begin] = cardinal cell of particle begin+q (particle_bins, unsorted).
// `rank_scratch` holds the per-particle slot claimed in its cell.
int bin_particles( const cb_grid& grid, const cb_positions& x, long long begin,
                   long long end, int* counts, unsigned* offsets, unsigned* permute,
                   int* cell_of, DeviceBuffer& rank_scratch, DeviceBuffer& scan_scratch,
                   cudaStream_t stream, int refine = 1 );

int launch_grid_for( long long work_items, int block );

// cb_verlet_build with the particle count left on the device (cb_verlet.cu): x->n is then the
// host-side BOUND every array is sized for, *n_dev (<= x->n, written earlier on the same
// stream) the real count; [begin, end) must lie below it.  v2 kernels only.
bool verlet_devcount_supported();
int verlet_build_devcount( cb_verlet* v, const cb_positions* x, long long begin, long long end,
                           double radius, double cell_size_ratio, const double* grid_min,
                           const double* grid_max, long long max_neigh, int algorithm,
                           int layout, int build_op, void* stream, const long long* n_dev );

} // namespace cb

// The LinkedCellList handle (shared by cb_lcl.cu and the LCL-direct traversal).
struct cb_lcl
{
    cb_grid grid;
    cb_grid stencil_grid;
    int cell_range = 1;
    int sorted = 0;
    long long begin = 0, end = 0;
    long long num_cells = 0;
    cb::DeviceBuffer counts, offsets, permute, bins, bins_alt, rank, scan, field_scratch;
    bool built = false;
};
