// Cloud-in-cell cell index and weights exactly as the reference computes them (auxPM.c:296-322, 580-600):
//   X = (double)Pos * (Nmesh/Box); I = (unsigned)X; D = X - I; T = 1 - D; DY, TY *= W; I >= Nmesh -> 0 for y, z.
#pragma once

namespace mgp {

struct Cic {
  unsigned ix, iy, iz;     // cell (y,z wrapped; x global)
  double dx, dy, dz, tx, ty, tz;
};

__device__ __forceinline__ Cic cic_of(const float4 p, double scale, unsigned N, double W) {
  Cic q;
  const double X = (double) p.x * scale, Y = (double) p.y * scale, Z = (double) p.z * scale;
  q.ix = (unsigned) X; q.iy = (unsigned) Y; q.iz = (unsigned) Z;
  q.dx = X - (double) q.ix; q.dy = Y - (double) q.iy; q.dz = Z - (double) q.iz;
  q.tx = 1.0 - q.dx; q.ty = 1.0 - q.dy; q.tz = 1.0 - q.dz;
  q.dy *= W; q.ty *= W;
  if (q.iy >= N) q.iy = 0;
  if (q.iz >= N) q.iz = 0;
  return q;
}

// (double) f without the conversion unit (64-bit conversions issue at a quarter of the FP64 rate): exact for every
// normal float; zero, denormals, infinities and NaNs take the real conversion.
__device__ __forceinline__ double f2d_exact(float f) {
  const unsigned b = __float_as_uint(f);
  const unsigned e = b & 0x7f800000u;
  if (e == 0u || e == 0x7f800000u) return (double) f;
  return __hiloint2double((int) ((b & 0x80000000u) | (((b & 0x7fffffffu) >> 3) + 0x38000000u)), (int) (b << 29));
}

// floor of 0 <= x < 2^31 as an integer and as a double, again without conversions: x + 2^52 rounded towards
// -infinity leaves floor(x) in the low mantissa word
__device__ __forceinline__ unsigned floor_u32(double x, double &as_double) {
  const double t = __dadd_rd(x, 4503599627370496.0);
  as_double = t - 4503599627370496.0;
  return (unsigned) __double2loint(t);
}

}  // namespace mgp
