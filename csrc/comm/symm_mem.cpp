// Symmetric-memory arena for --ddp-backend b200: physical allocations that every rank of one NVSwitch domain maps
// (cuMemCreate / cuMemExportToShareableHandle / cuMemImportFromShareableHandle / cuMemMap) plus an NVLS multicast
// alias (cuMulticastCreate / AddDevice / BindMem) for multimem.ld_reduce / multimem.st.
//
// The reference has no peer memory at all (unicore/distributed/utils.py:119-125 initialises NCCL and stops there); in
// round 1 this framework borrowed torch.distributed._symmetric_memory for the allocation and rendezvous.  This file is
// the in-repo replacement: the driver-API half lives here, the exchange of the POSIX file descriptors between the
// ranks (SCM_RIGHTS over abstract unix sockets whose names travel through the c10d store) is a few lines of Python in
// unicore_b200/parallel/symm_mem.py.  Lifetime: every mapping, imported handle and address range is owned by a
// SymmAllocation object and released in its destructor in the order unmap -> release -> free, which is safe whatever
// the state of the peers (a crashed peer's physical memory stays valid until the last importer drops it).
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <cuda.h>
#include <cuda_runtime.h>
#include <torch/extension.h>
#include <unistd.h>

#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "comm_api.h"

namespace {

// ---- driver entry points (no link-time dependency on libcuda) ----------------------------------------------------------
template <typename Fn>
Fn driver_fn(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<Fn>(p);
}

struct Driver {
  CUresult (*MemCreate)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
  CUresult (*MemRelease)(CUmemGenericAllocationHandle);
  CUresult (*MemGetAllocationGranularity)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
  CUresult (*MemExportToShareableHandle)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
  CUresult (*MemImportFromShareableHandle)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
  CUresult (*MemAddressReserve)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
  CUresult (*MemAddressFree)(CUdeviceptr, size_t);
  CUresult (*MemMap)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
  CUresult (*MemUnmap)(CUdeviceptr, size_t);
  CUresult (*MemSetAccess)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
  CUresult (*MulticastCreate)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
  CUresult (*MulticastAddDevice)(CUmemGenericAllocationHandle, CUdevice);
  CUresult (*MulticastBindMem)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t,
                               unsigned long long);
  CUresult (*MulticastUnbind)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);
  CUresult (*MulticastGetGranularity)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
  CUresult (*DeviceGetAttribute)(int*, CUdevice_attribute, CUdevice);
  CUresult (*GetErrorString)(CUresult, const char**);
  bool ok = false;
};

const Driver& drv() {
  static Driver d;
  static std::once_flag once;
  std::call_once(once, [] {
    cudaFree(nullptr);  // make sure the primary context exists
#define UB_LOAD(field, name) d.field = driver_fn<decltype(d.field)>(name)
    UB_LOAD(MemCreate, "cuMemCreate");
    UB_LOAD(MemRelease, "cuMemRelease");
    UB_LOAD(MemGetAllocationGranularity, "cuMemGetAllocationGranularity");
    UB_LOAD(MemExportToShareableHandle, "cuMemExportToShareableHandle");
    UB_LOAD(MemImportFromShareableHandle, "cuMemImportFromShareableHandle");
    UB_LOAD(MemAddressReserve, "cuMemAddressReserve");
    UB_LOAD(MemAddressFree, "cuMemAddressFree");
    UB_LOAD(MemMap, "cuMemMap");
    UB_LOAD(MemUnmap, "cuMemUnmap");
    UB_LOAD(MemSetAccess, "cuMemSetAccess");
    UB_LOAD(MulticastCreate, "cuMulticastCreate");
    UB_LOAD(MulticastAddDevice, "cuMulticastAddDevice");
    UB_LOAD(MulticastBindMem, "cuMulticastBindMem");
    UB_LOAD(MulticastUnbind, "cuMulticastUnbind");
    UB_LOAD(MulticastGetGranularity, "cuMulticastGetGranularity");
    UB_LOAD(DeviceGetAttribute, "cuDeviceGetAttribute");
    UB_LOAD(GetErrorString, "cuGetErrorString");
#undef UB_LOAD
    d.ok = d.MemCreate && d.MemRelease && d.MemGetAllocationGranularity && d.MemExportToShareableHandle &&
           d.MemImportFromShareableHandle && d.MemAddressReserve && d.MemAddressFree && d.MemMap && d.MemUnmap &&
           d.MemSetAccess && d.DeviceGetAttribute;
  });
  return d;
}

void check(CUresult r, const char* what) {
  if (r == CUDA_SUCCESS) return;
  const char* msg = nullptr;
  if (drv().GetErrorString) drv().GetErrorString(r, &msg);
  TORCH_CHECK(false, "symm_mem: ", what, " failed: ", msg ? msg : "unknown error", " (", (int)r, ")");
}

size_t round_up(size_t v, size_t g) { return (v + g - 1) / g * g; }

CUmemAllocationProp local_prop(int device) {
  CUmemAllocationProp prop{};
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

// One symmetric allocation as seen by ONE rank.
class SymmAllocation {
 public:
  SymmAllocation(int64_t bytes, int64_t device, int64_t rank, int64_t world, bool want_multicast)
      : device_((int)device), rank_((int)rank), world_((int)world) {
    TORCH_CHECK(drv().ok, "symm_mem: the CUDA driver does not export the virtual-memory API");
    TORCH_CHECK(world >= 1 && world <= ub::kMaxPeers && rank >= 0 && rank < world && bytes > 0);
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
    cudaFree(nullptr);
    const CUmemAllocationProp prop = local_prop(device_);
    size_t gran = 0;
    check(drv().MemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED), "cuMemGetAllocationGranularity");
    multicast_ok_ = want_multicast && multicast_supported(device_) && world_ > 1;
    if (multicast_ok_) {
      CUmulticastObjectProp mp = mc_prop(0);
      size_t mgran = 0;
      if (drv().MulticastGetGranularity(&mgran, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mgran > gran)
        gran = mgran;
    }
    size_ = round_up((size_t)bytes, gran);
    check(drv().MemCreate(&handle_, size_, &prop, 0), "cuMemCreate");
    have_handle_ = true;
    peer_handles_.assign(world_, 0);
    peer_imported_.assign(world_, false);
    ptrs_.assign(world_, 0);
    ptrs_[rank_] = map(handle_);
  }

  ~SymmAllocation() { release(); }

  static bool multicast_supported(int device) {
    if (!drv().ok || !drv().MulticastCreate || !drv().MulticastAddDevice || !drv().MulticastBindMem) return false;
    int v = 0;
    if (drv().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS) return false;
    return v != 0;
  }
  static bool supported(int device) {
    if (!drv().ok) return false;
    int v = 0;
    if (drv().DeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR_SUPPORTED, device) != CUDA_SUCCESS)
      return false;
    return v != 0;
  }

  int64_t size() const { return (int64_t)size_; }
  bool has_multicast() const { return mc_ptr_ != 0; }
  bool multicast_possible() const { return multicast_ok_; }

  // file descriptor of the local physical allocation; the caller sends it to the peers and closes it
  int64_t export_fd() {
    int fd = -1;
    check(drv().MemExportToShareableHandle(&fd, handle_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
          "cuMemExportToShareableHandle");
    return fd;
  }
  // map a peer's allocation (fd received over SCM_RIGHTS; closed here)
  void import_peer(int64_t peer, int64_t fd) {
    TORCH_CHECK(peer >= 0 && peer < world_ && peer != rank_ && !peer_imported_[peer]);
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
    CUmemGenericAllocationHandle h;
    const CUresult r = drv().MemImportFromShareableHandle(&h, reinterpret_cast<void*>((intptr_t)fd),
                                                          CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    ::close((int)fd);
    check(r, "cuMemImportFromShareableHandle");
    peer_handles_[peer] = h;
    peer_imported_[peer] = true;
    ptrs_[peer] = map(h);
  }

  // ---- NVLS: rank 0 creates the multicast object, everybody adds its device, binds its memory, maps the alias ----
  int64_t multicast_create() {
    TORCH_CHECK(multicast_ok_ && !have_mc_);
    CUmulticastObjectProp mp = mc_prop(size_);
    check(drv().MulticastCreate(&mc_handle_, &mp), "cuMulticastCreate");
    have_mc_ = true;
    int fd = -1;
    check(drv().MemExportToShareableHandle(&fd, mc_handle_, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
          "cuMemExportToShareableHandle(multicast)");
    return fd;
  }
  void multicast_import(int64_t fd) {
    TORCH_CHECK(multicast_ok_ && !have_mc_);
    const CUresult r = drv().MemImportFromShareableHandle(&mc_handle_, reinterpret_cast<void*>((intptr_t)fd),
                                                          CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    ::close((int)fd);
    check(r, "cuMemImportFromShareableHandle(multicast)");
    have_mc_ = true;
  }
  void multicast_add_device() {
    TORCH_CHECK(have_mc_);
    check(drv().MulticastAddDevice(mc_handle_, (CUdevice)device_), "cuMulticastAddDevice");
  }
  // (call after EVERY rank has added its device)
  void multicast_bind_and_map() {
    TORCH_CHECK(have_mc_ && mc_ptr_ == 0);
    const c10::cuda::CUDAGuard guard((c10::DeviceIndex)device_);
    check(drv().MulticastBindMem(mc_handle_, 0, handle_, 0, size_, 0), "cuMulticastBindMem");
    mc_bound_ = true;
    mc_ptr_ = map(mc_handle_);
  }

  std::vector<int64_t> ptrs() const { return std::vector<int64_t>(ptrs_.begin(), ptrs_.end()); }
  int64_t multicast_ptr() const { return (int64_t)mc_ptr_; }

  // the local mapping as a tensor (keeps `self` alive through the deleter)
  at::Tensor tensor(std::shared_ptr<SymmAllocation> self, int64_t numel, at::ScalarType dtype) {
    TORCH_CHECK(numel >= 0 && (size_t)numel * c10::elementSize(dtype) <= size_);
    auto opts = at::TensorOptions().dtype(dtype).device(at::kCUDA, (c10::DeviceIndex)device_);
    return at::from_blob(reinterpret_cast<void*>(ptrs_[rank_]), {numel}, [self](void*) mutable { self.reset(); }, opts);
  }

  void release() {
    if (released_) return;
    released_ = true;
    cudaDeviceSynchronize();  // no kernel of this process may still touch the mappings
    if (mc_ptr_ != 0) unmap(mc_ptr_);
    if (mc_bound_ && drv().MulticastUnbind) drv().MulticastUnbind(mc_handle_, (CUdevice)device_, 0, size_);
    if (have_mc_) drv().MemRelease(mc_handle_);
    for (int p = 0; p < world_; ++p) {
      if (ptrs_[p] != 0) unmap(ptrs_[p]);
      if (p != rank_ && peer_imported_[p]) drv().MemRelease(peer_handles_[p]);
    }
    if (have_handle_) drv().MemRelease(handle_);
    mc_ptr_ = 0;
  }

 private:
  CUmulticastObjectProp mc_prop(size_t size) const {
    CUmulticastObjectProp mp{};
    mp.numDevices = (unsigned)world_;
    mp.size = size;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    mp.flags = 0;
    return mp;
  }
  CUdeviceptr map(CUmemGenericAllocationHandle h) {
    CUdeviceptr va = 0;
    check(drv().MemAddressReserve(&va, size_, 0, 0, 0), "cuMemAddressReserve");
    CUresult r = drv().MemMap(va, size_, 0, h, 0);
    if (r != CUDA_SUCCESS) {
      drv().MemAddressFree(va, size_);
      check(r, "cuMemMap");
    }
    CUmemAccessDesc acc{};
    acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    acc.location.id = device_;
    acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
    r = drv().MemSetAccess(va, size_, &acc, 1);
    if (r != CUDA_SUCCESS) {
      drv().MemUnmap(va, size_);
      drv().MemAddressFree(va, size_);
      check(r, "cuMemSetAccess");
    }
    return va;
  }
  void unmap(CUdeviceptr va) {
    drv().MemUnmap(va, size_);
    drv().MemAddressFree(va, size_);
  }

  int device_, rank_, world_;
  size_t size_ = 0;
  CUmemGenericAllocationHandle handle_{}, mc_handle_{};
  bool have_handle_ = false, have_mc_ = false, mc_bound_ = false, multicast_ok_ = false, released_ = false;
  std::vector<CUmemGenericAllocationHandle> peer_handles_;
  std::vector<bool> peer_imported_;
  std::vector<CUdeviceptr> ptrs_;
  CUdeviceptr mc_ptr_ = 0;
};

}  // namespace

void register_symm_mem(pybind11::module_& m) {
  namespace py = pybind11;
  py::class_<SymmAllocation, std::shared_ptr<SymmAllocation>>(m, "SymmAllocation")
      .def(py::init<int64_t, int64_t, int64_t, int64_t, bool>(), py::arg("bytes"), py::arg("device"), py::arg("rank"),
           py::arg("world"), py::arg("want_multicast") = true)
      .def("size", &SymmAllocation::size)
      .def("export_fd", &SymmAllocation::export_fd)
      .def("import_peer", &SymmAllocation::import_peer)
      .def("multicast_possible", &SymmAllocation::multicast_possible)
      .def("multicast_create", &SymmAllocation::multicast_create)
      .def("multicast_import", &SymmAllocation::multicast_import)
      .def("multicast_add_device", &SymmAllocation::multicast_add_device)
      .def("multicast_bind_and_map", &SymmAllocation::multicast_bind_and_map)
      .def("has_multicast", &SymmAllocation::has_multicast)
      .def("ptrs", &SymmAllocation::ptrs)
      .def("multicast_ptr", &SymmAllocation::multicast_ptr)
      .def("release", &SymmAllocation::release)
      .def("tensor", [](std::shared_ptr<SymmAllocation> self, int64_t numel, py::object dtype) {
        return self->tensor(self, numel, torch::python::detail::py_object_to_dtype(dtype));
      });
  m.def("symm_mem_supported", [](int64_t device) { return SymmAllocation::supported((int)device); });
  m.def("symm_multicast_supported", [](int64_t device) { return SymmAllocation::multicast_supported((int)device); });
}
