/* oracle/_ref run-time stand-ins -- TEST INFRASTRUCTURE ONLY.
 * libcore/tls.cpp is built on Boost.MultiIndex, which this image lacks; this file implements the same interface
 * (include/mitsuba/core/tls.h: detail::ThreadLocalBase and the four TLS life-cycle hooks) with a mutex-protected map
 * per object.  No arithmetic of the hot path lives here. */
#include <mitsuba/mitsuba.h>
#include <mitsuba/core/tls.h>
#include <map>
#include <mutex>
#include <thread>

MTS_NAMESPACE_BEGIN
namespace detail {
struct ThreadLocalBase::ThreadLocalPrivate {
    ConstructFunctor construct;
    DestructFunctor destruct;
    std::mutex mutex;
    std::map<std::thread::id, void *> data;
};
ThreadLocalBase::ThreadLocalBase(const ConstructFunctor &c, const DestructFunctor &d_) : d(new ThreadLocalPrivate()) {
    d->construct = c; d->destruct = d_;
}
ThreadLocalBase::~ThreadLocalBase() {
    for (auto &kv : d->data) d->destruct(kv.second);
}
void *ThreadLocalBase::get(bool &existed) {
    std::lock_guard<std::mutex> guard(d->mutex);
    auto it = d->data.find(std::this_thread::get_id());
    if (it != d->data.end()) { existed = true; return it->second; }
    existed = false;
    void *p = d->construct();
    d->data[std::this_thread::get_id()] = p;
    return p;
}
const void *ThreadLocalBase::get(bool &existed) const { return const_cast<ThreadLocalBase *>(this)->get(existed); }
void *ThreadLocalBase::get() { bool e; return get(e); }
const void *ThreadLocalBase::get() const { bool e; return get(e); }
void initializeGlobalTLS() {}
void destroyGlobalTLS() {}
void initializeLocalTLS() {}
void destroyLocalTLS() {}
}
MTS_NAMESPACE_END
