// Library plumbing: error text, device/memory helpers and the host-side CartesianGrid /
// LinkedCellStencil helpers of the C ABI (same inline arithmetic as the kernels).
#include <math.h>
#include <string.h>

#include <atomic>
#include <string>

#include "cb_common.cuh"
#include "cb_internal.h"

namespace cb
{

static thread_local std::string g_last_error = "";

static std::atomic<long long> g_launches{ 0 };
void note_launch() { g_launches.fetch_add( 1, std::memory_order_relaxed ); }

void set_error( const std::string& msg ) { g_last_error = msg; }

int fail( cb_status code, const char* what )
{
    g_last_error = what ? what : "";
    return (int)code;
}

int cuda_fail( cudaError_t err, const char* what, const char* file, int line )
{
    char buf[512];
    snprintf( buf, sizeof( buf ), "CUDA error %d (%s) in %s at %s:%d", (int)err,
              cudaGetErrorString( err ), what ? what : "?", file, line );
    g_last_error = buf;
    (void)cudaGetLastError(); // clear the sticky-free error
    return (int)CB_ERR_CUDA;
}

} // namespace cb

using namespace cb;

extern "C" int cb_version( void ) { return CB_VERSION_MAJOR * 1000 + CB_VERSION_MINOR; }

extern "C" const char* cb_last_error_string( void ) { return g_last_error.c_str(); }

extern "C" int cb_device_count( void )
{
    int n = 0;
    if ( cudaGetDeviceCount( &n ) != cudaSuccess )
    {
        (void)cudaGetLastError();
        return 0;
    }
    return n;
}

extern "C" int64_t cb_kernel_launch_count( void )
{
    return (int64_t)g_launches.load( std::memory_order_relaxed );
}

extern "C" int cb_set_device( int device )
{
    CB_CUDA( cudaSetDevice( device ) );
    return CB_OK;
}

extern "C" int cb_malloc( void** ptr, size_t bytes )
{
    if ( !ptr )
        return fail( CB_ERR_INVALID, "cb_malloc: null argument" );
    CB_CUDA( cudaMalloc( ptr, bytes ? bytes : 1 ) );
    return CB_OK;
}
extern "C" int cb_free( void* ptr )
{
    if ( ptr )
        CB_CUDA( cudaFree( ptr ) );
    return CB_OK;
}
extern "C" int cb_malloc_host( void** ptr, size_t bytes )
{
    if ( !ptr )
        return fail( CB_ERR_INVALID, "cb_malloc_host: null argument" );
    CB_CUDA( cudaMallocHost( ptr, bytes ? bytes : 1 ) );
    return CB_OK;
}
extern "C" int cb_free_host( void* ptr )
{
    if ( ptr )
        CB_CUDA( cudaFreeHost( ptr ) );
    return CB_OK;
}
extern "C" int cb_memcpy_h2d( void* dst, const void* src, size_t bytes, cb_stream_t s )
{
    if ( bytes )
        CB_CUDA( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyHostToDevice,
                                  (cudaStream_t)s ) );
    return CB_OK;
}
extern "C" int cb_memcpy_d2h( void* dst, const void* src, size_t bytes, cb_stream_t s )
{
    if ( bytes )
        CB_CUDA( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToHost,
                                  (cudaStream_t)s ) );
    return CB_OK;
}
extern "C" int cb_memcpy_d2d( void* dst, const void* src, size_t bytes, cb_stream_t s )
{
    if ( bytes )
        CB_CUDA( cudaMemcpyAsync( dst, src, bytes, cudaMemcpyDeviceToDevice,
                                  (cudaStream_t)s ) );
    return CB_OK;
}
extern "C" int cb_memset( void* dst, int value, size_t bytes, cb_stream_t s )
{
    if ( bytes )
        CB_CUDA( cudaMemsetAsync( dst, value, bytes, (cudaStream_t)s ) );
    return CB_OK;
}
extern "C" int cb_stream_synchronize( cb_stream_t s )
{
    CB_CUDA( cudaStreamSynchronize( (cudaStream_t)s ) );
    return CB_OK;
}

// ---- CartesianGrid (core/src/impl/Cabana_CartesianGrid.hpp) ----------------------------
extern "C" int cb_grid_init( cb_grid* g, const double* mn, const double* mx,
                             const double* delta )
{
    if ( !g || !mn || !mx || !delta )
        return fail( CB_ERR_INVALID, "cb_grid_init: null argument" );
    for ( int d = 0; d < 3; ++d ) // init :61-73
    {
        g->min[d] = mn[d];
        g->max[d] = mx[d];
        g->nx[d] = cells_between( mx[d], mn[d], 1.0 / delta[d] );
        g->dx[d] = ( mx[d] - mn[d] ) / g->nx[d];
        g->rdx[d] = 1.0 / g->dx[d];
    }
    return CB_OK;
}

extern "C" int cb_grid_locate_point( const cb_grid* g, const double* p, int32_t* ijk )
{
    if ( !g || !p || !ijk )
        return fail( CB_ERR_INVALID, "cb_grid_locate_point: null argument" );
    const Grid gg = to_grid( *g );
    for ( int d = 0; d < 3; ++d )
        ijk[d] = locate_1d( gg, d, p[d] );
    return CB_OK;
}

extern "C" double cb_grid_min_distance_to_point( const cb_grid* g, const double* x,
                                                 const int32_t* ijk )
{
    const Grid gg = to_grid( *g );
    return min_distance_sq( gg, x[0], x[1], x[2], ijk[0], ijk[1], ijk[2] );
}

extern "C" int cb_grid_cardinal_cell_index( const cb_grid* g, int i, int j, int k )
{
    return ( i * g->nx[1] + j ) * g->nx[2] + k;
}

extern "C" int cb_grid_ijk_bin_index( const cb_grid* g, int cardinal, int32_t* ijk )
{
    if ( !g || !ijk )
        return fail( CB_ERR_INVALID, "cb_grid_ijk_bin_index: null argument" );
    ijk[0] = cardinal / ( g->nx[1] * g->nx[2] );
    ijk[1] = ( cardinal / g->nx[2] ) % g->nx[1];
    ijk[2] = cardinal % g->nx[2];
    return CB_OK;
}

// ---- LinkedCellStencil (core/src/Cabana_LinkedCellList.hpp:55-66, :105-119) ------------
extern "C" int cb_stencil_cell_range( double cell_size_ratio )
{
    return (int)ceil( 1 / cell_size_ratio ); // :63
}

extern "C" int cb_stencil_get_cells( const cb_grid* sg, int cell_range, int cell,
                                     int32_t* mn, int32_t* mx )
{
    if ( !sg || !mn || !mx )
        return fail( CB_ERR_INVALID, "cb_stencil_get_cells: null argument" );
    int32_t ijk[3];
    cb_grid_ijk_bin_index( sg, cell, ijk );
    for ( int d = 0; d < 3; ++d )
    {
        mn[d] = ( ijk[d] - cell_range > 0 ) ? ijk[d] - cell_range : 0;
        mx[d] = ( ijk[d] + cell_range + 1 < sg->nx[d] ) ? ijk[d] + cell_range + 1
                                                        : sg->nx[d];
    }
    return CB_OK;
}
