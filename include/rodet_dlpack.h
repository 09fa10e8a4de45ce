/*
 * Minimal declaration of the DLPack tensor structs (DLPack v0.8 ABI, the layout
 * `torch.utils.dlpack.to_dlpack` produces in a "dltensor" capsule) used by the
 * rod_dl_* entry points of rodet_b200.h.  Written from the DLPack specification;
 * field order and widths must not change.
 */
#ifndef RODET_DLPACK_H_
#define RODET_DLPACK_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum { ROD_kDLCPU = 1, ROD_kDLCUDA = 2, ROD_kDLCUDAHost = 3, ROD_kDLCUDAManaged = 13 };
enum { ROD_kDLInt = 0, ROD_kDLUInt = 1, ROD_kDLFloat = 2 };

typedef struct DLDevice {
  int32_t device_type;
  int32_t device_id;
} DLDevice;

typedef struct DLDataType {
  uint8_t code;
  uint8_t bits;
  uint16_t lanes;
} DLDataType;

typedef struct DLTensor {
  void* data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t* shape;
  int64_t* strides;   /* in elements; NULL = compact row-major */
  uint64_t byte_offset;
} DLTensor;

typedef struct DLManagedTensor {
  DLTensor dl_tensor;  /* first member: a DLManagedTensor* is a valid DLTensor* */
  void* manager_ctx;
  void (*deleter)(struct DLManagedTensor* self);
} DLManagedTensor;

#ifdef __cplusplus
}
#endif
#endif
