#pragma once
// Stand-in for the Boost.Thread subset libcore/{thread,lock,sched}.cpp use, over the C++ standard library.
#include <thread>
#include <mutex>
#include <condition_variable>
#include <chrono>
#include <stdexcept>
namespace boost {
namespace posix_time {
    typedef std::chrono::system_clock::time_point ptime;
    inline std::chrono::milliseconds milliseconds(long ms) { return std::chrono::milliseconds(ms); }
}
inline posix_time::ptime get_system_time() { return std::chrono::system_clock::now(); }
struct thread_resource_error : std::runtime_error { thread_resource_error() : std::runtime_error("thread_resource_error") {} };
struct thread_interrupted {};
template <class M> struct with_scoped_lock : M { typedef std::unique_lock<M> scoped_lock; };
typedef with_scoped_lock<std::mutex> mutex;
typedef with_scoped_lock<std::recursive_mutex> recursive_mutex;
struct timed_mutex : std::timed_mutex { typedef std::unique_lock<timed_mutex> scoped_lock;
    bool timed_lock(const posix_time::ptime &t) { return try_lock_until(t); } };
struct recursive_timed_mutex : std::recursive_timed_mutex { typedef std::unique_lock<recursive_timed_mutex> scoped_lock;
    bool timed_lock(const posix_time::ptime &t) { return try_lock_until(t); } };
template <class M> using lock_guard = std::lock_guard<M>;
template <class M> using unique_lock = std::unique_lock<M>;
class condition_variable_any : public std::condition_variable_any {
public:
    template <class L> bool timed_wait(L &lock, const posix_time::ptime &t) { return wait_until(lock, t) == std::cv_status::no_timeout; }
};
class thread : public std::thread {
public:
    thread() {}
    template <class F, class... A> explicit thread(F &&f, A &&... a) : std::thread(std::forward<F>(f), std::forward<A>(a)...) {}
    thread(thread &&o) : std::thread(std::move(static_cast<std::thread &>(o))) {}
    thread &operator=(thread &&o) { std::thread::operator=(std::move(static_cast<std::thread &>(o))); return *this; }
    static unsigned hardware_concurrency() { return std::thread::hardware_concurrency(); }
};
namespace this_thread {
    inline void yield() { std::this_thread::yield(); }
    template <class D> void sleep(const D &d) { std::this_thread::sleep_for(d); }
}
}
