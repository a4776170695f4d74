/*
 * lmz_dlpack.h -- the subset of the DLPack tensor-exchange ABI this library reads.
 *
 * DLPack is a public, frozen C struct layout; PyTorch's `tensor.__dlpack__()`
 * hands out a PyCapsule named "dltensor" holding a `DLManagedTensor*` laid out
 * exactly as below.  The shim only ever BORROWS these structs for the duration
 * of a call: it never calls `deleter` and never renames the capsule.
 */
#ifndef LMZ_DLPACK_H
#define LMZ_DLPACK_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifndef DLPACK_VERSION   /* do not clash with a real dlpack.h if one is included first */

typedef enum {
  kDLCPU = 1,
  kDLCUDA = 2,
  kDLCUDAHost = 3,
  kDLCUDAManaged = 13
} DLDeviceType;

typedef struct {
  int32_t device_type;   /* DLDeviceType */
  int32_t device_id;
} DLDevice;

typedef enum {
  kDLInt = 0,
  kDLUInt = 1,
  kDLFloat = 2,
  kDLBfloat = 4,
  kDLBool = 6
} DLDataTypeCode;

typedef struct {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DLDataType;

typedef struct {
  void *data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t *shape;
  int64_t *strides;      /* NULL => compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;
  void *manager_ctx;
  void (*deleter)(struct DLManagedTensor *self);
} DLManagedTensor;

#endif /* DLPACK_VERSION */

#ifdef __cplusplus
}
#endif
#endif
