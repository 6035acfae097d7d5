// device_array.hpp -- drop-in for the reference's buffer ABI (ThirdParty/pcl_gpu_containers/include/
// device_array.h:57-258, kernel_containers.h:54-102, src/device_memory.cpp:108-322): ref-counted pitched device
// buffers with the same names, namespaces and member functions, so code written against the reference
// (src/visodo.cpp, src/keyframe_align.cpp) compiles unchanged.  Header-only; written from scratch on the CUDA
// runtime.  Differences: errors throw std::runtime_error instead of calling exit(0)
// (ThirdParty/pcl_gpu_containers/src/error.cpp:42-46), and copies are issued on the legacy default stream only
// when no stream is given; every copy is followed by cudaStreamSynchronize(0) as in the reference
// (ThirdParty/pcl_gpu_containers/src/device_memory.cpp:284-307): a pageable-memory cudaMemcpy may return before the DMA
// has landed and a device-to-device one is asynchronous, and the kernels run on a non-blocking stream.
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pcl {
namespace gpu {

inline void cudaSafeCallImpl(cudaError_t e, const char* file, int line)
{
  if (e != cudaSuccess)
    throw std::runtime_error(std::string(cudaGetErrorString(e)) + " at " + file + ":" + std::to_string(line));
}
#define cudaSafeCall(expr) ::pcl::gpu::cudaSafeCallImpl((expr), __FILE__, __LINE__)

static inline int divUp(int total, int grain) { return (total + grain - 1) / grain; }

// ---- kernel-side views (kernel_containers.h) ------------------------------------------------------------
template <typename T> struct DevPtr {
  typedef T elem_type;
  T* data;
  DevPtr() : data(0) {}
  DevPtr(T* d) : data(d) {}
  size_t elemSize() const { return sizeof(T); }
  operator T*() { return data; }
  operator const T*() const { return data; }
};

template <typename T> struct PtrSz : public DevPtr<T> {
  PtrSz() : size(0) {}
  PtrSz(T* d, size_t s) : DevPtr<T>(d), size(s) {}
  size_t size;
};

template <typename T> struct PtrStep : public DevPtr<T> {
  PtrStep() : step(0) {}
  PtrStep(T* d, size_t s) : DevPtr<T>(d), step(s) {}
  size_t step;  // bytes
  T* ptr(int y = 0) { return (T*)((char*)DevPtr<T>::data + y * step); }
  const T* ptr(int y = 0) const { return (const T*)((const char*)DevPtr<T>::data + y * step); }
};

template <typename T> struct PtrStepSz : public PtrStep<T> {
  PtrStepSz() : cols(0), rows(0) {}
  PtrStepSz(int r, int c, T* d, size_t s) : PtrStep<T>(d, s), cols(c), rows(r) {}
  int cols, rows;
};

// ---- reference-counted storage ------------------------------------------------------------------------------
class DeviceMemory {
 public:
  DeviceMemory() : data_(0), sizeBytes_(0), refcount_(0) {}
  explicit DeviceMemory(size_t bytes) : data_(0), sizeBytes_(0), refcount_(0) { create(bytes); }
  DeviceMemory(void* ptr, size_t bytes) : data_(ptr), sizeBytes_(bytes), refcount_(0) {}  // user memory, not owned
  DeviceMemory(const DeviceMemory& o) : data_(o.data_), sizeBytes_(o.sizeBytes_), refcount_(o.refcount_) { if (refcount_) ++*refcount_; }
  ~DeviceMemory() { release(); }
  DeviceMemory& operator=(const DeviceMemory& o)
  {
    if (this != &o) {
      if (o.refcount_) ++*o.refcount_;
      release();
      data_ = o.data_; sizeBytes_ = o.sizeBytes_; refcount_ = o.refcount_;
    }
    return *this;
  }
  void create(size_t bytes)
  {
    if (bytes == sizeBytes_) return;
    if (bytes > 0) {
      if (data_) release();
      sizeBytes_ = bytes;
      cudaSafeCall(cudaMalloc(&data_, sizeBytes_));
      refcount_ = new int(1);
    }
  }
  void release()
  {
    if (refcount_ && --*refcount_ == 0) { delete refcount_; cudaFree(data_); }
    data_ = 0; sizeBytes_ = 0; refcount_ = 0;
  }
  void copyTo(DeviceMemory& other) const
  {
    if (empty()) { other.release(); return; }
    other.create(sizeBytes_);
    cudaSafeCall(cudaMemcpy(other.data_, data_, sizeBytes_, cudaMemcpyDeviceToDevice)); cudaSafeCall(cudaStreamSynchronize(0));
  }
  void upload(const void* host, size_t bytes) { create(bytes); cudaSafeCall(cudaMemcpy(data_, host, bytes, cudaMemcpyHostToDevice)); cudaSafeCall(cudaStreamSynchronize(0)); }
  void download(void* host) const { cudaSafeCall(cudaMemcpy(host, data_, sizeBytes_, cudaMemcpyDeviceToHost)); cudaSafeCall(cudaStreamSynchronize(0)); }
  void swap(DeviceMemory& o) { std::swap(data_, o.data_); std::swap(sizeBytes_, o.sizeBytes_); std::swap(refcount_, o.refcount_); }
  template <class T> T* ptr() { return (T*)data_; }
  template <class T> const T* ptr() const { return (const T*)data_; }
  bool empty() const { return !data_; }
  size_t sizeBytes() const { return sizeBytes_; }

 private:
  void* data_;
  size_t sizeBytes_;
  int* refcount_;
};

class DeviceMemory2D {
 public:
  DeviceMemory2D() : data_(0), step_(0), colsBytes_(0), rows_(0), refcount_(0) {}
  DeviceMemory2D(int rows, int colsBytes) : data_(0), step_(0), colsBytes_(0), rows_(0), refcount_(0) { create(rows, colsBytes); }
  DeviceMemory2D(int rows, int colsBytes, void* data, size_t step)
      : data_(data), step_(step), colsBytes_(colsBytes), rows_(rows), refcount_(0) {}
  DeviceMemory2D(const DeviceMemory2D& o)
      : data_(o.data_), step_(o.step_), colsBytes_(o.colsBytes_), rows_(o.rows_), refcount_(o.refcount_) { if (refcount_) ++*refcount_; }
  ~DeviceMemory2D() { release(); }
  DeviceMemory2D& operator=(const DeviceMemory2D& o)
  {
    if (this != &o) {
      if (o.refcount_) ++*o.refcount_;
      release();
      data_ = o.data_; step_ = o.step_; colsBytes_ = o.colsBytes_; rows_ = o.rows_; refcount_ = o.refcount_;
    }
    return *this;
  }
  void create(int rows, int colsBytes)
  {
    if (rows_ == rows && colsBytes_ == colsBytes) return;
    if (rows > 0 && colsBytes > 0) {
      if (data_) release();
      colsBytes_ = colsBytes; rows_ = rows;
      cudaSafeCall(cudaMallocPitch(&data_, &step_, colsBytes_, rows_));
      refcount_ = new int(1);
    }
  }
  void release()
  {
    if (refcount_ && --*refcount_ == 0) { delete refcount_; cudaFree(data_); }
    data_ = 0; step_ = 0; colsBytes_ = 0; rows_ = 0; refcount_ = 0;
  }
  void copyTo(DeviceMemory2D& other) const
  {
    if (empty()) { other.release(); return; }
    other.create(rows_, colsBytes_);
    cudaSafeCall(cudaMemcpy2D(other.data_, other.step_, data_, step_, colsBytes_, rows_, cudaMemcpyDeviceToDevice)); cudaSafeCall(cudaStreamSynchronize(0));
  }
  void upload(const void* host, size_t hostStep, int rows, int colsBytes)
  {
    create(rows, colsBytes);
    cudaSafeCall(cudaMemcpy2D(data_, step_, host, hostStep, colsBytes_, rows_, cudaMemcpyHostToDevice)); cudaSafeCall(cudaStreamSynchronize(0));
  }
  void download(void* host, size_t hostStep) const
  {
    cudaSafeCall(cudaMemcpy2D(host, hostStep, data_, step_, colsBytes_, rows_, cudaMemcpyDeviceToHost)); cudaSafeCall(cudaStreamSynchronize(0));
  }
  void swap(DeviceMemory2D& o)
  {
    std::swap(data_, o.data_); std::swap(step_, o.step_); std::swap(colsBytes_, o.colsBytes_);
    std::swap(rows_, o.rows_); std::swap(refcount_, o.refcount_);
  }
  template <class T> T* ptr(int y = 0) { return (T*)((char*)data_ + y * step_); }
  template <class T> const T* ptr(int y = 0) const { return (const T*)((const char*)data_ + y * step_); }
  bool empty() const { return !data_; }
  int colsBytes() const { return colsBytes_; }
  int rows() const { return rows_; }
  size_t step() const { return step_; }

 private:
  void* data_;
  size_t step_;
  int colsBytes_, rows_;
  int* refcount_;
};

// ---- typed containers (device_array.h) -------------------------------------------------------------------------
template <class T> class DeviceArray : public DeviceMemory {
 public:
  typedef T type;
  enum { elem_size = sizeof(T) };
  DeviceArray() {}
  explicit DeviceArray(size_t size) : DeviceMemory(size * elem_size) {}
  DeviceArray(T* ptr, size_t size) : DeviceMemory(ptr, size * elem_size) {}
  void create(size_t size) { DeviceMemory::create(size * elem_size); }
  void copyTo(DeviceArray& other) const { DeviceMemory::copyTo(other); }
  void upload(const T* host, size_t size) { DeviceMemory::upload(host, size * elem_size); }
  void download(T* host) const { DeviceMemory::download(host); }
  void upload(const std::vector<T>& data) { upload(data.data(), data.size()); }
  void download(std::vector<T>& data) const { data.resize(size()); if (!data.empty()) download(data.data()); }
  void swap(DeviceArray& o) { DeviceMemory::swap(o); }
  T* ptr() { return DeviceMemory::ptr<T>(); }
  const T* ptr() const { return DeviceMemory::ptr<T>(); }
  operator T*() { return ptr(); }
  operator const T*() const { return ptr(); }
  size_t size() const { return sizeBytes() / elem_size; }
  operator PtrSz<T>() const { return PtrSz<T>(const_cast<T*>(ptr()), size()); }
};

template <class T> class DeviceArray2D : public DeviceMemory2D {
 public:
  typedef T type;
  enum { elem_size = sizeof(T) };
  DeviceArray2D() {}
  DeviceArray2D(int rows, int cols) : DeviceMemory2D(rows, cols * elem_size) {}
  DeviceArray2D(int rows, int cols, void* data, size_t stepBytes) : DeviceMemory2D(rows, cols * elem_size, data, stepBytes) {}
  void create(int rows, int cols) { DeviceMemory2D::create(rows, cols * elem_size); }
  void copyTo(DeviceArray2D& other) const { DeviceMemory2D::copyTo(other); }
  void upload(const void* host, size_t hostStep, int rows, int cols) { DeviceMemory2D::upload(host, hostStep, rows, cols * elem_size); }
  void download(void* host, size_t hostStep) const { DeviceMemory2D::download(host, hostStep); }
  void upload(const std::vector<T>& data, int cols) { upload(data.data(), cols * elem_size, (int)(data.size() / cols), cols); }
  void download(std::vector<T>& data, int& elem_step) const
  {
    elem_step = cols();
    data.resize((size_t)cols() * rows());
    if (!data.empty()) download(data.data(), colsBytes());
  }
  void swap(DeviceArray2D& o) { DeviceMemory2D::swap(o); }
  T* ptr(int y = 0) { return DeviceMemory2D::ptr<T>(y); }
  const T* ptr(int y = 0) const { return DeviceMemory2D::ptr<T>(y); }
  operator T*() { return ptr(); }
  operator const T*() const { return ptr(); }
  int cols() const { return colsBytes() / elem_size; }
  int rows() const { return DeviceMemory2D::rows(); }
  size_t elem_step() const { return step() / elem_size; }
  operator PtrStep<T>() const { return PtrStep<T>(const_cast<T*>(ptr()), step()); }
  operator PtrStepSz<T>() const { return PtrStepSz<T>(rows(), cols(), const_cast<T*>(ptr()), step()); }
};

}  // namespace gpu
namespace device {
using pcl::gpu::PtrStep;
using pcl::gpu::PtrStepSz;
using pcl::gpu::PtrSz;
}  // namespace device
}  // namespace pcl
