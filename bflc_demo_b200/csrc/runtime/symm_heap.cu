// See symm_heap.hpp.  Host-only code (driver API resolved at run time so the library
// still loads on a box without libcuda, e.g. the CPU authoring container).
#include "symm_heap.hpp"

#include <cuda.h>
#include <cuda_runtime.h>
#include <cstring>
#include <stdexcept>
#include <sys/socket.h>
#include <sys/syscall.h>
#include <sys/un.h>
#include <unistd.h>

namespace bflc {

namespace {

template <typename Fn>
Fn drv(const char* name) {
  void* sym = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &sym, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<Fn>(sym);
}

#define DRV(name) static auto p_##name = drv<decltype(&name)>(#name)

void rt_check(cudaError_t e, const char* what) {
  if (e != cudaSuccess)
    throw std::runtime_error(std::string("SymmHeap: ") + what + ": " + cudaGetErrorString(e));
}
void cu_check(CUresult r, const char* what) {
  if (r != CUDA_SUCCESS)
    throw std::runtime_error(std::string("SymmHeap: ") + what + ": CUresult " +
                             std::to_string(static_cast<int>(r)));
}

struct FdBlob {
  int32_t pid;
  int32_t fd;
  uint64_t bytes;
};

// Duplicate `fd` of process `pid` into this process (Linux >= 5.6).
int steal_fd(int pid, int fd) {
  if (pid == getpid()) return dup(fd);
#ifndef SYS_pidfd_open
#define SYS_pidfd_open 434
#endif
#ifndef SYS_pidfd_getfd
#define SYS_pidfd_getfd 438
#endif
  int pidfd = static_cast<int>(syscall(SYS_pidfd_open, pid, 0));
  if (pidfd < 0) return -1;
  int got = static_cast<int>(syscall(SYS_pidfd_getfd, pidfd, fd, 0));
  close(pidfd);
  return got;
}

size_t round_up(size_t x, size_t g) { return (x + g - 1) / g * g; }

// --- SCM_RIGHTS helpers -------------------------------------------------------------------
socklen_t abstract_addr(const std::string& name, sockaddr_un* a) {
  std::memset(a, 0, sizeof(*a));
  a->sun_family = AF_UNIX;
  const size_t n = name.size() < sizeof(a->sun_path) - 2 ? name.size() : sizeof(a->sun_path) - 2;
  std::memcpy(a->sun_path + 1, name.data(), n);  // leading NUL = abstract namespace
  return static_cast<socklen_t>(offsetof(sockaddr_un, sun_path) + 1 + n);
}
bool send_fd(int sock, int fd) {
  char tag = fd >= 0 ? 'F' : 'N';
  iovec iov{&tag, 1};
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  if (fd >= 0) {
    std::memset(ctl, 0, sizeof(ctl));
    msg.msg_control = ctl;
    msg.msg_controllen = sizeof(ctl);
    cmsghdr* c = CMSG_FIRSTHDR(&msg);
    c->cmsg_level = SOL_SOCKET;
    c->cmsg_type = SCM_RIGHTS;
    c->cmsg_len = CMSG_LEN(sizeof(int));
    std::memcpy(CMSG_DATA(c), &fd, sizeof(int));
  }
  return sendmsg(sock, &msg, 0) == 1;
}
int recv_fd(int sock) {  // -1: no fd sent, -2: error
  char tag = 0;
  iovec iov{&tag, 1};
  msghdr msg{};
  msg.msg_iov = &iov;
  msg.msg_iovlen = 1;
  alignas(cmsghdr) char ctl[CMSG_SPACE(sizeof(int))];
  std::memset(ctl, 0, sizeof(ctl));
  msg.msg_control = ctl;
  msg.msg_controllen = sizeof(ctl);
  if (recvmsg(sock, &msg, 0) != 1) return -2;
  if (tag != 'F') return -1;
  for (cmsghdr* c = CMSG_FIRSTHDR(&msg); c; c = CMSG_NXTHDR(&msg, c))
    if (c->cmsg_level == SOL_SOCKET && c->cmsg_type == SCM_RIGHTS) {
      int fd;
      std::memcpy(&fd, CMSG_DATA(c), sizeof(int));
      return fd;
    }
  return -2;
}

CUmemAllocationProp vmm_prop(int device) {
  CUmemAllocationProp prop;
  std::memset(&prop, 0, sizeof(prop));
  prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  prop.location.id = device;
  prop.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return prop;
}

void* map_handle(CUmemGenericAllocationHandle h, size_t bytes, size_t gran, int device) {
  DRV(cuMemAddressReserve);
  DRV(cuMemMap);
  DRV(cuMemSetAccess);
  CUdeviceptr va = 0;
  cu_check(p_cuMemAddressReserve(&va, bytes, gran, 0, 0), "cuMemAddressReserve");
  cu_check(p_cuMemMap(va, bytes, 0, h, 0), "cuMemMap");
  CUmemAccessDesc acc;
  std::memset(&acc, 0, sizeof(acc));
  acc.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  acc.location.id = device;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  cu_check(p_cuMemSetAccess(va, bytes, &acc, 1), "cuMemSetAccess");
  return reinterpret_cast<void*>(va);
}

}  // namespace

bool SymmHeap::multicast_supported(int device) {
  DRV(cuDeviceGetAttribute);
  DRV(cuMulticastCreate);
  if (!p_cuDeviceGetAttribute || !p_cuMulticastCreate) return false;
  int v = 0;
  if (p_cuDeviceGetAttribute(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, device) != CUDA_SUCCESS)
    return false;
  return v != 0;
}

SymmHeap::SymmHeap(size_t bytes, int rank, int world, int device, Mode mode)
    : rank_(rank), world_(world), device_(device), mode_(mode) {
  rt_check(cudaSetDevice(device), "cudaSetDevice");
  rt_check(cudaFree(nullptr), "context init");
  peers_.assign(static_cast<size_t>(world), nullptr);
  if (mode_ == Mode::VMM) {
    DRV(cuMemGetAllocationGranularity);
    DRV(cuMemCreate);
    DRV(cuMemExportToShareableHandle);
    if (!p_cuMemCreate || !p_cuMemGetAllocationGranularity || !p_cuMemExportToShareableHandle)
      throw std::runtime_error("SymmHeap: VMM driver entry points unavailable");
    CUmemAllocationProp prop = vmm_prop(device);
    size_t gran = 0;
    cu_check(p_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
             "cuMemGetAllocationGranularity");
    if (gran < (size_t(2) << 20)) gran = size_t(2) << 20;
    bytes_ = round_up(bytes, gran);
    CUmemGenericAllocationHandle h;
    cu_check(p_cuMemCreate(&h, bytes_, &prop, 0), "cuMemCreate");
    mem_handle_ = h;
    int fd = -1;
    cu_check(p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0),
             "cuMemExportToShareableHandle");
    export_fd_ = fd;
    local_ = map_handle(h, bytes_, gran, device);
    peer_handles_.assign(static_cast<size_t>(world), 0);
  } else {
    bytes_ = round_up(bytes, size_t(2) << 20);
    rt_check(cudaMalloc(&local_, bytes_), "cudaMalloc");
  }
  rt_check(cudaMemset(local_, 0, bytes_), "cudaMemset");
  rt_check(cudaDeviceSynchronize(), "sync");
  peers_[static_cast<size_t>(rank)] = local_;
}

SymmHeap::~SymmHeap() {
  cudaDeviceSynchronize();
  if (mode_ == Mode::IPC) {
    for (int r = 0; r < world_; ++r)
      if (r != rank_ && peers_[static_cast<size_t>(r)])
        cudaIpcCloseMemHandle(peers_[static_cast<size_t>(r)]);
    cudaFree(local_);
  } else if (mode_ == Mode::LOCAL) {
    cudaFree(local_);
  } else {
    DRV(cuMemUnmap);
    DRV(cuMemAddressFree);
    DRV(cuMemRelease);
    if (p_cuMemUnmap && p_cuMemAddressFree && p_cuMemRelease) {
      if (mc_va_) {
        p_cuMemUnmap(reinterpret_cast<CUdeviceptr>(mc_va_), bytes_);
        p_cuMemAddressFree(reinterpret_cast<CUdeviceptr>(mc_va_), bytes_);
      }
      if (mc_handle_) p_cuMemRelease(mc_handle_);
      for (int r = 0; r < world_; ++r) {
        void* p = peers_[static_cast<size_t>(r)];
        if (!p) continue;
        p_cuMemUnmap(reinterpret_cast<CUdeviceptr>(p), bytes_);
        p_cuMemAddressFree(reinterpret_cast<CUdeviceptr>(p), bytes_);
        if (r != rank_ && peer_handles_[static_cast<size_t>(r)])
          p_cuMemRelease(peer_handles_[static_cast<size_t>(r)]);
      }
      if (mem_handle_) p_cuMemRelease(mem_handle_);
    }
    if (export_fd_ >= 0) close(export_fd_);
    if (mc_fd_ >= 0) close(mc_fd_);
  }
  if (listen_fd_ >= 0) close(listen_fd_);
}

std::string SymmHeap::export_handle() const {
  if (mode_ == Mode::IPC) {
    cudaIpcMemHandle_t h;
    rt_check(cudaIpcGetMemHandle(&h, local_), "cudaIpcGetMemHandle");
    return std::string(reinterpret_cast<const char*>(&h), sizeof(h));
  }
  if (mode_ == Mode::VMM) {
    FdBlob b{static_cast<int32_t>(getpid()), export_fd_, bytes_};
    return std::string(reinterpret_cast<const char*>(&b), sizeof(b));
  }
  return std::string();
}

void SymmHeap::import_handles(const std::vector<std::string>& blobs) {
  if (mode_ == Mode::LOCAL) return;
  if (static_cast<int>(blobs.size()) != world_)
    throw std::runtime_error("SymmHeap: import_handles needs one blob per rank");
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    const std::string& blob = blobs[static_cast<size_t>(r)];
    if (mode_ == Mode::IPC) {
      if (blob.size() != sizeof(cudaIpcMemHandle_t))
        throw std::runtime_error("SymmHeap: bad IPC handle blob");
      cudaIpcMemHandle_t h;
      std::memcpy(&h, blob.data(), sizeof(h));
      void* p = nullptr;
      rt_check(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess),
               "cudaIpcOpenMemHandle");
      peers_[static_cast<size_t>(r)] = p;
    } else {
      DRV(cuMemImportFromShareableHandle);
      DRV(cuMemGetAllocationGranularity);
      if (blob.size() != sizeof(FdBlob)) throw std::runtime_error("SymmHeap: bad fd blob");
      FdBlob b;
      std::memcpy(&b, blob.data(), sizeof(b));
      if (b.bytes != bytes_) throw std::runtime_error("SymmHeap: asymmetric heap sizes");
      int fd = steal_fd(b.pid, b.fd);
      if (fd < 0) throw std::runtime_error("SymmHeap: pidfd_getfd failed (ptrace scope?)");
      CUmemGenericAllocationHandle h;
      cu_check(p_cuMemImportFromShareableHandle(
                   &h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                   CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
               "cuMemImportFromShareableHandle");
      close(fd);
      peer_handles_[static_cast<size_t>(r)] = h;
      CUmemAllocationProp prop = vmm_prop(device_);
      size_t gran = 0;
      cu_check(p_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
               "granularity");
      peers_[static_cast<size_t>(r)] = map_handle(h, bytes_, gran, device_);
    }
  }
}

std::string SymmHeap::fd_listen(const std::string& unique_tag) {
  if (listen_fd_ >= 0) close(listen_fd_);
  listen_fd_ = socket(AF_UNIX, SOCK_STREAM, 0);
  if (listen_fd_ < 0) throw std::runtime_error("SymmHeap: socket() failed");
  const std::string name = "bflc_symm_" + unique_tag + "_" + std::to_string(rank_);
  sockaddr_un a;
  const socklen_t len = abstract_addr(name, &a);
  if (bind(listen_fd_, reinterpret_cast<sockaddr*>(&a), len) != 0 || listen(listen_fd_, 64) != 0)
    throw std::runtime_error("SymmHeap: bind/listen on abstract socket failed");
  return name;
}

// All-to-all exchange of one fd per rank (my_fd < 0 = nothing to share).  connect() completes
// against the listen backlog, so [connect all] -> [accept all + send] -> [recv all] cannot
// deadlock as long as every rank listened before the caller's barrier.
std::vector<int> SymmHeap::exchange_fds(const std::vector<std::string>& names, int my_fd) {
  if (listen_fd_ < 0) throw std::runtime_error("SymmHeap: fd_listen() was not called");
  std::vector<int> out_sock(static_cast<size_t>(world_), -1), got(static_cast<size_t>(world_), -1);
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    int s = socket(AF_UNIX, SOCK_STREAM, 0);
    sockaddr_un a;
    const socklen_t len = abstract_addr(names[static_cast<size_t>(r)], &a);
    if (s < 0 || connect(s, reinterpret_cast<sockaddr*>(&a), len) != 0)
      throw std::runtime_error("SymmHeap: connect to peer socket failed");
    out_sock[static_cast<size_t>(r)] = s;
  }
  for (int i = 0; i < world_ - 1; ++i) {
    int c = accept(listen_fd_, nullptr, nullptr);
    if (c < 0 || !send_fd(c, my_fd)) throw std::runtime_error("SymmHeap: accept/send_fd failed");
    close(c);
  }
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    const int fd = recv_fd(out_sock[static_cast<size_t>(r)]);
    close(out_sock[static_cast<size_t>(r)]);
    if (fd == -2) throw std::runtime_error("SymmHeap: recv_fd failed");
    got[static_cast<size_t>(r)] = fd;
  }
  return got;
}

void SymmHeap::map_peer_fd(int r, int fd) {
  DRV(cuMemImportFromShareableHandle);
  DRV(cuMemGetAllocationGranularity);
  CUmemGenericAllocationHandle h;
  cu_check(p_cuMemImportFromShareableHandle(&h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
                                            CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR),
           "cuMemImportFromShareableHandle");
  close(fd);
  peer_handles_[static_cast<size_t>(r)] = h;
  CUmemAllocationProp prop = vmm_prop(device_);
  size_t gran = 0;
  cu_check(p_cuMemGetAllocationGranularity(&gran, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED),
           "granularity");
  peers_[static_cast<size_t>(r)] = map_handle(h, bytes_, gran, device_);
}

void SymmHeap::import_via_sockets(const std::vector<std::string>& names) {
  if (mode_ != Mode::VMM) throw std::runtime_error("SymmHeap: socket import needs VMM mode");
  std::vector<int> fds = exchange_fds(names, export_fd_);
  for (int r = 0; r < world_; ++r) {
    if (r == rank_) continue;
    if (fds[static_cast<size_t>(r)] < 0) throw std::runtime_error("SymmHeap: peer sent no fd");
    map_peer_fd(r, fds[static_cast<size_t>(r)]);
  }
}

bool SymmHeap::mc_import_via_sockets(const std::vector<std::string>& names) {
  if (mode_ != Mode::VMM) return false;
  DRV(cuMemImportFromShareableHandle);
  DRV(cuMulticastAddDevice);
  std::vector<int> fds = exchange_fds(names, rank_ == 0 ? mc_fd_ : -1);
  if (rank_ != 0) {
    const int fd = fds[0];
    if (fd < 0) { err_ = "rank 0 has no multicast object: " + err_; return false; }
    CUmemGenericAllocationHandle h;
    CUresult r = p_cuMemImportFromShareableHandle(
        &h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) {
      err_ = "import multicast handle: CUresult " + std::to_string(static_cast<int>(r));
      return false;
    }
    mc_handle_ = h;
  } else if (!mc_handle_) {
    return false;
  }
  CUresult r = p_cuMulticastAddDevice(mc_handle_, device_);
  if (r != CUDA_SUCCESS) {
    err_ = "cuMulticastAddDevice: CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

std::string SymmHeap::mc_create_and_export() {
  if (mode_ != Mode::VMM || !multicast_supported(device_)) return std::string();
  DRV(cuMulticastCreate);
  DRV(cuMulticastGetGranularity);
  DRV(cuMemExportToShareableHandle);
  CUmulticastObjectProp mp;
  std::memset(&mp, 0, sizeof(mp));
  mp.numDevices = static_cast<unsigned>(world_);
  mp.size = bytes_;
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  size_t g = 0;
  if (p_cuMulticastGetGranularity(&g, &mp, CU_MULTICAST_GRANULARITY_MINIMUM) != CUDA_SUCCESS ||
      (g && bytes_ % g != 0)) {
    err_ = "multicast granularity mismatch";
    return std::string();
  }
  CUmemGenericAllocationHandle h;
  CUresult r = p_cuMulticastCreate(&h, &mp);
  if (r != CUDA_SUCCESS) {
    err_ = "cuMulticastCreate: CUresult " + std::to_string(static_cast<int>(r));
    return std::string();
  }
  mc_handle_ = h;
  int fd = -1;
  r = p_cuMemExportToShareableHandle(&fd, h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (r != CUDA_SUCCESS) {
    err_ = "export multicast handle: CUresult " + std::to_string(static_cast<int>(r));
    return std::string();
  }
  mc_fd_ = fd;
  FdBlob b{static_cast<int32_t>(getpid()), fd, bytes_};
  return std::string(reinterpret_cast<const char*>(&b), sizeof(b));
}

bool SymmHeap::mc_import_and_add(const std::string& blob) {
  if (mode_ != Mode::VMM || blob.size() != sizeof(FdBlob)) return false;
  DRV(cuMemImportFromShareableHandle);
  DRV(cuMulticastAddDevice);
  if (!mc_handle_) {
    FdBlob b;
    std::memcpy(&b, blob.data(), sizeof(b));
    int fd = steal_fd(b.pid, b.fd);
    if (fd < 0) { err_ = "pidfd_getfd(multicast) failed"; return false; }
    CUmemGenericAllocationHandle h;
    CUresult r = p_cuMemImportFromShareableHandle(
        &h, reinterpret_cast<void*>(static_cast<intptr_t>(fd)),
        CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
    close(fd);
    if (r != CUDA_SUCCESS) {
      err_ = "import multicast handle: CUresult " + std::to_string(static_cast<int>(r));
      return false;
    }
    mc_handle_ = h;
  }
  CUresult r = p_cuMulticastAddDevice(mc_handle_, device_);
  if (r != CUDA_SUCCESS) {
    err_ = "cuMulticastAddDevice: CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  return true;
}

bool SymmHeap::mc_bind_and_map() {
  if (mode_ != Mode::VMM || !mc_handle_) return false;
  DRV(cuMulticastBindMem);
  DRV(cuMulticastGetGranularity);
  CUresult r = p_cuMulticastBindMem(mc_handle_, 0, mem_handle_, 0, bytes_, 0);
  if (r != CUDA_SUCCESS) {
    err_ = "cuMulticastBindMem: CUresult " + std::to_string(static_cast<int>(r));
    return false;
  }
  mc_bound_ = true;
  try {
    CUmulticastObjectProp mp;
    std::memset(&mp, 0, sizeof(mp));
    mp.numDevices = static_cast<unsigned>(world_);
    mp.size = bytes_;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t g = size_t(2) << 20;
    p_cuMulticastGetGranularity(&g, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED);
    mc_va_ = map_handle(mc_handle_, bytes_, g, device_);
  } catch (const std::exception& e) {
    err_ = e.what();
    mc_va_ = nullptr;
    return false;
  }
  return true;
}

}  // namespace bflc
