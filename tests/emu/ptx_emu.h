// ptx_emu.h -- TEST INFRASTRUCTURE ONLY: host models of the PTX helpers of nm_device.cuh (included by it under
// NM_HOST_EMU, after cuda_emu.h).  Shared-memory addresses are byte offsets into the emulated window nm::smem.
#pragma once
#include <atomic>
#include <chrono>
#include <mutex>
#include <thread>
#ifndef EMU_DEADLOCK_SECONDS
#define EMU_DEADLOCK_SECONDS 90
#endif

namespace nm {
extern float smem[];                                   // the window (defined by the emulator translation unit)

inline uint32_t smem_u32(const void* p) { return uint32_t(reinterpret_cast<const char*>(p) - reinterpret_cast<const char*>(smem)); }
inline char* smem_at(uint32_t off) { return reinterpret_cast<char*>(smem) + off; }

// mbarrier: a phase completes when the expected arrivals have arrived AND the transaction count is back to zero
struct EmuBar {
  std::mutex m;
  int count = 0, pending = 0;
  long tx = 0;
  std::atomic<uint32_t> phase{0};
  void settle() { if (pending <= 0 && tx == 0) { pending = count; phase.store(phase.load() ^ 1u); } }
};
inline EmuBar g_emu_bars[64 * 1024 / 2];               // keyed by (smem offset / 8)
inline EmuBar& emu_bar(const uint64_t* bar) { return g_emu_bars[smem_u32(bar) >> 3]; }

inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuBar& b = emu_bar(bar);
  std::lock_guard<std::mutex> l(b.m);
  b.count = int(count); b.pending = int(count); b.tx = 0; b.phase.store(0);
}
inline void mbar_fence_init() {}
inline void mbar_arrive(uint64_t* bar) {
  EmuBar& b = emu_bar(bar);
  std::lock_guard<std::mutex> l(b.m);
  --b.pending; b.settle();
}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {            // mbarrier.arrive.expect_tx
  EmuBar& b = emu_bar(bar);
  std::lock_guard<std::mutex> l(b.m);
  b.tx += long(bytes); --b.pending; b.settle();
}
inline void mbar_complete_tx(uint64_t* bar, uint32_t bytes) {
  EmuBar& b = emu_bar(bar);
  std::lock_guard<std::mutex> l(b.m);
  b.tx -= long(bytes); b.settle();
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {                // try_wait.parity: done once the phase of that parity completed
  EmuBar& b = emu_bar(bar);
  const auto t0 = std::chrono::steady_clock::now();
  unsigned long spins = 0;
  while (b.phase.load(std::memory_order_acquire) == (parity & 1u)) {
    std::this_thread::yield();
    if ((++spins & 0xFFFF) == 0 && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(EMU_DEADLOCK_SECONDS)) {
      // deadlock detector: say who waits on what, then stop the process
      std::lock_guard<std::mutex> l(b.m);
      fprintf(stderr, "[emu] DEADLOCK? block %u thread %u waits on mbarrier at smem offset %u for parity %u: phase %u pending %d/%d tx %ld\n",
              blockIdx.x, threadIdx.x, smem_u32(bar), parity & 1u, b.phase.load(), b.pending, b.count, b.tx);
      std::this_thread::sleep_for(std::chrono::milliseconds(300));      // let the other waiters report too
      std::_Exit(3);
    }
  }
  std::lock_guard<std::mutex> l(b.m);                                  // (orders the producer's writes before the consumer's reads)
}
inline void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  std::memcpy(dst_smem, src_gmem, bytes);
  mbar_complete_tx(bar, bytes);
}
inline void named_bar_sync64(int id) { emu::named_barrier(id, 64)->arrive_and_wait(); }
}  // namespace nm
