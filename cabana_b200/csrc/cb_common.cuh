// Shared host/device helpers for the cabana_b200 kernels (sm_100a).
//
// The arithmetic on every in/out decision path is spelled with __d*_rn intrinsics on
// the device (never contracted to FMA) and plain operators on the host (this file is
// compiled with -Xcompiler -ffp-contract=off), so host and device agree bit for bit
// with the reference's un-contracted source semantics.
#pragma once

#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/cabana_b200.h"

#define CB_HD __host__ __device__ __forceinline__
#define CB_D __device__ __forceinline__

#if defined( __CUDA_ARCH__ )
#define CB_MUL( a, b ) __dmul_rn( ( a ), ( b ) )
#define CB_ADD( a, b ) __dadd_rn( ( a ), ( b ) )
#define CB_SUB( a, b ) __dsub_rn( ( a ), ( b ) )
#else
#define CB_MUL( a, b ) ( ( a ) * ( b ) )
#define CB_ADD( a, b ) ( ( a ) + ( b ) )
#define CB_SUB( a, b ) ( ( a ) - ( b ) )
#endif

namespace cb
{

constexpr unsigned kFullMask = 0xffffffffu;
constexpr int kNumSMs = 148; // B200

// ---------------------------------------------------------------------------------
// Position / field accessors (cabana_b200.h: cb_positions, cb_field).
// ---------------------------------------------------------------------------------
struct PosAccess
{
    const double* base;
    long long outer_stride;
    long long comp_stride;
    int vlen;
    int vshift; // log2(vlen) when vlen is a power of two, else -1

    CB_HD long long offset( long long i ) const
    {
        if ( vlen == 1 )
            return outer_stride * i;
        if ( vshift >= 0 )
            return outer_stride * ( i >> vshift ) + ( i & ( vlen - 1 ) );
        return outer_stride * ( i / vlen ) + ( i % vlen );
    }
    CB_HD double get( long long i, int d ) const
    {
        return base[offset( i ) + comp_stride * d];
    }
};

inline int log2_if_pow2( int v )
{
    if ( v <= 0 || ( v & ( v - 1 ) ) )
        return -1;
    int s = 0;
    while ( ( 1 << s ) < v )
        ++s;
    return s;
}

inline PosAccess make_access( const cb_positions& x )
{
    PosAccess a;
    a.base = x.base;
    a.outer_stride = x.outer_stride;
    a.comp_stride = x.comp_stride;
    a.vlen = x.vlen;
    a.vshift = log2_if_pow2( x.vlen );
    return a;
}

struct FieldAccess
{
    char* base;
    long long outer_stride; // elements
    long long comp_stride;  // elements
    int vlen;
    int vshift;
    int num_comp;
    int elem_bytes;

    CB_HD long long offset( long long i ) const
    {
        if ( vlen == 1 )
            return outer_stride * i;
        if ( vshift >= 0 )
            return outer_stride * ( i >> vshift ) + ( i & ( vlen - 1 ) );
        return outer_stride * ( i / vlen ) + ( i % vlen );
    }
};

inline FieldAccess make_access( const cb_field& f )
{
    FieldAccess a;
    a.base = (char*)f.base;
    a.outer_stride = f.outer_stride;
    a.comp_stride = f.comp_stride;
    a.vlen = f.vlen;
    a.vshift = log2_if_pow2( f.vlen );
    a.num_comp = f.num_comp;
    a.elem_bytes = f.elem_bytes;
    return a;
}

// ---------------------------------------------------------------------------------
// CartesianGrid (core/src/impl/Cabana_CartesianGrid.hpp), exact arithmetic.
// ---------------------------------------------------------------------------------
struct Grid
{
    double min[3];
    double max[3];
    double dx[3];
    double rdx[3];
    int nx[3];
};

inline Grid to_grid( const cb_grid& g )
{
    Grid r;
    for ( int d = 0; d < 3; ++d )
    {
        r.min[d] = g.min[d];
        r.max[d] = g.max[d];
        r.dx[d] = g.dx[d];
        r.rdx[d] = g.rdx[d];
        r.nx[d] = g.nx[d];
    }
    return r;
}

// cellsBetween (:279-283): floor((max - min) * rdelta), sub and mul rounded separately.
CB_HD int cells_between( double hi, double lo, double rdelta )
{
#if defined( __CUDA_ARCH__ )
    return __double2int_rd( __dmul_rn( __dsub_rn( hi, lo ), rdelta ) );
#else
    return (int)floor( ( hi - lo ) * rdelta );
#endif
}

// locatePoint (:171-182): only the exact upper edge is clamped.
CB_HD int locate_1d( const Grid& g, int d, double p )
{
    int c = cells_between( p, g.min[d], g.rdx[d] );
    return ( c == g.nx[d] ) ? c - 1 : c;
}

// Sub-cell index on the m-times refined internal grid: user cell `cu` (exact, from
// locate_1d) times m plus the sub-cell the point falls in.  Monotone non-decreasing in p.
CB_HD int fine_index_1d( const Grid& g, int d, double p, int cu, int m )
{
    const double t = CB_MUL( CB_SUB( p, g.min[d] ), g.rdx[d] );
    int sub = (int)( ( t - (double)cu ) * (double)m );
    sub = sub < 0 ? 0 : ( sub > m - 1 ? m - 1 : sub );
    return cu * m + sub;
}

// One dimension of minDistanceToPoint (:212-220): returns rx*rx for cell index c.
CB_HD double min_dist_1d_sq( const Grid& g, int d, double x, int c )
{
    double xc = CB_ADD( g.min[d], CB_MUL( (double)c + 0.5, g.dx[d] ) );
    double rx = CB_SUB( fabs( CB_SUB( x, xc ) ), CB_MUL( 0.5, g.dx[d] ) );
    rx = ( rx > 0.0 ) ? rx : 0.0;
    return CB_MUL( rx, rx );
}

// minDistanceToPoint (:207-223): ((0 + rx^2) + ry^2) + rz^2.
CB_HD double min_distance_sq( const Grid& g, double x, double y, double z, int i,
                              int j, int k )
{
    double r = min_dist_1d_sq( g, 0, x, i ); // 0.0 + v == v for v >= 0
    r = CB_ADD( r, min_dist_1d_sq( g, 1, y, j ) );
    r = CB_ADD( r, min_dist_1d_sq( g, 2, z, k ) );
    return r;
}

CB_HD int cardinal_index( const Grid& g, int i, int j, int k )
{
    return ( i * g.nx[1] + j ) * g.nx[2] + k;
}

// Pair distance (core/src/Cabana_VerletList.hpp:464-469): d = xp - xn; s = ((0 + d0^2) +
// d1^2) + d2^2 with every product and sum rounded separately.
CB_HD double pair_dist_sq( double xp, double yp, double zp, double xn, double yn,
                           double zn )
{
    double d0 = CB_SUB( xp, xn );
    double d1 = CB_SUB( yp, yn );
    double d2 = CB_SUB( zp, zn );
    double s = CB_MUL( d0, d0 );
    s = CB_ADD( s, CB_MUL( d1, d1 ) );
    s = CB_ADD( s, CB_MUL( d2, d2 ) );
    return s;
}

// NeighborDiscriminator<HalfNeighborTag>::isValid coordinate part
// (core/src/Cabana_NeighborList.hpp:139-149).
CB_HD bool half_criterion( double xp, double yp, double zp, double xn, double yn,
                           double zn )
{
    return ( ( xn > xp ) ||
             ( ( xn == xp ) &&
               ( ( yn > yp ) || ( ( yn == yp ) && ( zn > zp ) ) ) ) );
}

#if defined( __CUDACC__ )
CB_D unsigned lane_id() { return threadIdx.x & 31u; }
CB_D unsigned lanemask_lt()
{
    unsigned m;
    asm( "mov.u32 %0, %%lanemask_lt;" : "=r"( m ) );
    return m;
}
CB_D int warp_reduce_sum( int v )
{
    return __reduce_add_sync( kFullMask, v );
}
CB_D int warp_reduce_max( int v )
{
    return __reduce_max_sync( kFullMask, v );
}
CB_D double warp_reduce_sum( double v )
{
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 )
        v += __shfl_xor_sync( kFullMask, v, o );
    return v;
}
CB_D long long warp_reduce_sum( long long v )
{
#pragma unroll
    for ( int o = 16; o > 0; o >>= 1 )
        v += __shfl_xor_sync( kFullMask, v, o );
    return v;
}
#endif

} // namespace cb
