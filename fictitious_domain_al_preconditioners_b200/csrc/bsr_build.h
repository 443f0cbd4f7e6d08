// bsr_build.h — CSR -> BSR conversion on the device (setup phase of fdal_finalize; csrc/bsr_build.cu).
#pragma once
#include <cuda_runtime.h>

namespace fdal {

struct BsrBuilt {
  int *brp = nullptr;     // [nbr + 1] block-row pointers
  int *bcj = nullptr;     // [nblk] block columns, ascending inside a block row
  double *bv = nullptr;   // [nblk * b * b] blocks stored contiguously, row-major inside a block
  long long nblk = 0;
};

enum { BSR_BUILD_OK = 0, BSR_BUILD_DECLINED = 1, BSR_BUILD_CUDA_ERROR = -1 };

// Device arrays in (scalar CSR with 32-bit row pointers, nr a multiple of b), device arrays out (cudaMalloc'ed
// here; the caller owns them).  max_row_entries: the largest number of scalar entries in one block row (the b
// consecutive CSR rows), which sizes the per-warp hash set.  DECLINED — nothing allocated — when blocking would
// store more than max_fill x the scalar non-zeros, the block count would overflow 32 bits, or a block row is too
// long for the shared-memory hash set; the caller then keeps the host conversion (csrc/host_finalize.h), whose
// result this routine reproduces bit for bit on matrices without duplicate (row, column) entries.
int bsr_from_csr_device(cudaStream_t stream, int sms, int nr, long long nnz, const int *rp, const int *ci, const double *v,
                        int b, int max_row_entries, double max_fill, BsrBuilt *out);

}  // namespace fdal
