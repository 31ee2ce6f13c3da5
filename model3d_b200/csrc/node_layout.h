// Byte stride of one wide-BVH node record in HBM, shared by the builders (wide_bvh.h) and the
// traversal (trace_core.cuh).  The node itself is 80 bytes (five 128-bit loads that straddle three
// or four 32-byte sectors).  M3D_NODE_BYTES=96 pads every node to a sector boundary and fetches it
// with two 256-bit loads and one 128-bit load (three sector lookups per lane).  Measured on B200
// (C2, same rays, same hits): 2.235 ms vs 2.225 ms for the 80-byte stride -- the 20 % larger node
// array costs what the fewer L1TEX lookups save, so 80 stays the default.
#pragma once
#ifndef M3D_NODE_BYTES
#define M3D_NODE_BYTES 80
#endif
#if M3D_NODE_BYTES != 80 && M3D_NODE_BYTES != 96
#error "M3D_NODE_BYTES must be 80 or 96"
#endif
#define M3D_NODE_QUADS (M3D_NODE_BYTES / 16)
