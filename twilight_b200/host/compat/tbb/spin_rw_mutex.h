// host/compat: tbb::spin_rw_mutex stand-in backed by std::shared_mutex. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <atomic>
#include <shared_mutex>

namespace tbb {
class spin_rw_mutex {
  public:
    class scoped_lock {
      public:
        scoped_lock(spin_rw_mutex &m, bool write = true) : m_(m), write_(write) {
            if (write_) m_.mtx_.lock();
            else m_.mtx_.lock_shared();
        }
        ~scoped_lock() {
            if (write_) m_.mtx_.unlock();
            else m_.mtx_.unlock_shared();
        }
        scoped_lock(const scoped_lock &) = delete;
        scoped_lock &operator=(const scoped_lock &) = delete;

      private:
        spin_rw_mutex &m_;
        bool write_;
    };

  private:
    std::shared_mutex mtx_;
};
} // namespace tbb
