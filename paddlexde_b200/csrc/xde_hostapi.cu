// xde_hostapi.cu -- the torch-free host surface of libxde_b200: device memory, copies, streams, a DLPack producer
// and the one collective of the path.  With these a host layer needs nothing but ctypes + numpy
// (paddlexde_b200/_native.py); PyTorch / Paddle users keep handing their own tensors in as raw pointers.
//
// Reference protocol this serves: the reference's tensors are paddle.Tensor objects created by the caller and by
// eager ops inside the solver loop (solver/base_adaptive_solver.py:25-31, base_fixed_solver.py:119-143); here the
// solver's outputs are caller-visible device buffers exported through DLPack (paddle.utils.dlpack.from_dlpack /
// torch.from_dlpack / numpy via a host copy).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>

#include "xde_common.cuh"

namespace xde {

// ---- DLPack (v0.8 ABI: struct layouts restated from the public dlpack.h) ----
struct DLDevice {
  int32_t device_type;  // kDLCUDA = 2
  int32_t device_id;
};
struct DLDataType {
  uint8_t code;  // kDLInt = 0, kDLUInt = 1, kDLFloat = 2
  uint8_t bits;
  uint16_t lanes;
};
struct DLTensor {
  void *data;
  DLDevice device;
  int32_t ndim;
  DLDataType dtype;
  int64_t *shape;
  int64_t *strides;
  uint64_t byte_offset;
};
struct DLManagedTensor {
  DLTensor dl_tensor;
  void *manager_ctx;
  void (*deleter)(DLManagedTensor *self);
};

struct DlCtx {
  int64_t shape[8];
  int owns;  // free the device buffer with the capsule
};

static void dl_deleter(DLManagedTensor *self) {
  if (!self) return;
  DlCtx *c = static_cast<DlCtx *>(self->manager_ctx);
  if (c && c->owns && self->dl_tensor.data) cudaFreeAsync(self->dl_tensor.data, 0);
  delete c;
  delete self;
}

}  // namespace xde

using namespace xde;

extern "C" {

XDE_EXPORT int xde_device_count(int32_t *n) {
  XDE_REQUIRE(n, XDE_E_BAD_ARG, "null argument");
  int c = 0;
  XDE_CUDA_CHECK(cudaGetDeviceCount(&c));
  *n = c;
  return XDE_OK;
}
XDE_EXPORT int xde_set_device(int32_t dev) {
  XDE_CUDA_CHECK(cudaSetDevice(dev));
  return XDE_OK;
}
XDE_EXPORT int xde_get_device(int32_t *dev) {
  XDE_REQUIRE(dev, XDE_E_BAD_ARG, "null argument");
  int d = 0;
  XDE_CUDA_CHECK(cudaGetDevice(&d));
  *dev = d;
  return XDE_OK;
}
// stream-ordered allocation from the device's default memory pool (kept warm: see scratch_alloc)
XDE_EXPORT int xde_malloc(void **ptr, uint64_t bytes, void *stream) {
  XDE_REQUIRE(ptr, XDE_E_BAD_ARG, "null argument");
  *ptr = nullptr;
  if (bytes == 0) return XDE_OK;
  XDE_CUDA_CHECK(scratch_alloc(ptr, (size_t)bytes, (cudaStream_t)stream));
  return XDE_OK;
}
XDE_EXPORT int xde_free(void *ptr, void *stream) {
  if (ptr) XDE_CUDA_CHECK(cudaFreeAsync(ptr, (cudaStream_t)stream));
  return XDE_OK;
}
// kind: 0 host->device, 1 device->host, 2 device->device.  Asynchronous with respect to the host only when the host
// buffer is pinned (xde_host_alloc); pageable buffers are staged by the driver.
XDE_EXPORT int xde_memcpy_async(void *dst, const void *src, uint64_t bytes, int32_t kind, void *stream) {
  XDE_REQUIRE((dst && src) || bytes == 0, XDE_E_BAD_ARG, "null argument");
  XDE_REQUIRE(kind >= 0 && kind <= 2, XDE_E_BAD_ARG, "kind must be 0 (H2D), 1 (D2H) or 2 (D2D)");
  if (bytes == 0) return XDE_OK;
  const cudaMemcpyKind k = kind == 0 ? cudaMemcpyHostToDevice : (kind == 1 ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice);
  XDE_CUDA_CHECK(cudaMemcpyAsync(dst, src, (size_t)bytes, k, (cudaStream_t)stream));
  return XDE_OK;
}
XDE_EXPORT int xde_memset_async(void *dst, int32_t value, uint64_t bytes, void *stream) {
  XDE_REQUIRE(dst || bytes == 0, XDE_E_BAD_ARG, "null argument");
  if (bytes) XDE_CUDA_CHECK(cudaMemsetAsync(dst, value, (size_t)bytes, (cudaStream_t)stream));
  return XDE_OK;
}
XDE_EXPORT int xde_host_alloc(void **ptr, uint64_t bytes) {  // pinned host memory
  XDE_REQUIRE(ptr, XDE_E_BAD_ARG, "null argument");
  XDE_CUDA_CHECK(cudaMallocHost(ptr, (size_t)bytes));
  return XDE_OK;
}
XDE_EXPORT int xde_host_free(void *ptr) {
  if (ptr) XDE_CUDA_CHECK(cudaFreeHost(ptr));
  return XDE_OK;
}
XDE_EXPORT int xde_stream_create(void **stream) {
  XDE_REQUIRE(stream, XDE_E_BAD_ARG, "null argument");
  cudaStream_t s;
  XDE_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = (void *)s;
  return XDE_OK;
}
XDE_EXPORT int xde_stream_destroy(void *stream) {
  if (stream) XDE_CUDA_CHECK(cudaStreamDestroy((cudaStream_t)stream));
  return XDE_OK;
}
XDE_EXPORT int xde_stream_synchronize(void *stream) {
  XDE_CUDA_CHECK(cudaStreamSynchronize((cudaStream_t)stream));
  return XDE_OK;
}

// DLPack producer: a DLManagedTensor* describing a contiguous device buffer (dtype_code 2 = float, 0 = int; bits 32 /
// 64).  owns != 0: the capsule's deleter frees the buffer (xde_free on the default stream); 0: the buffer is borrowed
// and must outlive every consumer.  The caller wraps the pointer in a PyCapsule named "dltensor".
XDE_EXPORT void *xde_dlpack_wrap(void *dev_ptr, int32_t ndim, const int64_t *shape, int32_t dtype_code, int32_t bits,
                                 int32_t device_id, int32_t owns) {
  if (ndim < 0 || ndim > 8 || (ndim > 0 && !shape)) {
    set_last_error("xde_dlpack_wrap: ndim must be 0..8");
    return nullptr;
  }
  DLManagedTensor *m = new DLManagedTensor();
  DlCtx *c = new DlCtx();
  for (int i = 0; i < ndim; ++i) c->shape[i] = shape[i];
  c->owns = owns;
  m->dl_tensor.data = dev_ptr;
  m->dl_tensor.device.device_type = 2;  // kDLCUDA
  m->dl_tensor.device.device_id = device_id;
  m->dl_tensor.ndim = ndim;
  m->dl_tensor.dtype.code = (uint8_t)dtype_code;
  m->dl_tensor.dtype.bits = (uint8_t)bits;
  m->dl_tensor.dtype.lanes = 1;
  m->dl_tensor.shape = c->shape;
  m->dl_tensor.strides = nullptr;  // compact row-major
  m->dl_tensor.byte_offset = 0;
  m->manager_ctx = c;
  m->deleter = dl_deleter;
  return m;
}
// for a capsule that was never consumed
XDE_EXPORT void xde_dlpack_release(void *managed) {
  DLManagedTensor *m = static_cast<DLManagedTensor *>(managed);
  if (m && m->deleter) m->deleter(m);
}

// the ONLY collective of the path -- the sum of the adjoint parameter gradients over the batch shards
// (the DataParallel gradient all-reduce of example/D3STN/train_dde.py:201-202,454-456).  comm: an ncclComm_t created by
// the caller (PaddlePaddle's / PyTorch's / its own); in place on `buf` (n fp32 values), ordered on `stream`.  NCCL is
// resolved at the first call (dlopen of the process's libnccl.so.2): the library itself has no link-time dependency.
XDE_EXPORT int xde_allreduce_grads(void *comm, float *buf, int64_t n, void *stream) {
  XDE_REQUIRE(comm && buf && n >= 0, XDE_E_BAD_ARG, "null argument");
  typedef int (*allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
  typedef const char *(*errstr_fn)(int);
  static allreduce_fn fn = nullptr;
  static errstr_fn es = nullptr;
  if (!fn) {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    XDE_REQUIRE(h, XDE_E_CUDA, "xde_allreduce_grads: libnccl.so.2 cannot be loaded (%s)", dlerror());
    fn = (allreduce_fn)dlsym(h, "ncclAllReduce");
    es = (errstr_fn)dlsym(h, "ncclGetErrorString");
    XDE_REQUIRE(fn, XDE_E_CUDA, "xde_allreduce_grads: ncclAllReduce not found in libnccl");
  }
  if (n == 0) return XDE_OK;
  const int rc = fn(buf, buf, (size_t)n, /*ncclFloat32*/ 7, /*ncclSum*/ 0, comm, (cudaStream_t)stream);
  XDE_REQUIRE(rc == 0, XDE_E_CUDA, "ncclAllReduce failed: %s", es ? es(rc) : "?");
  return XDE_OK;
}

}  // extern "C"
