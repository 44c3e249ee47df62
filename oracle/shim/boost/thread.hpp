// Oracle tooling only: stand-in for boost/thread.hpp on top of <thread>/<mutex>.
#ifndef ORACLE_SHIM_BOOST_THREAD_HPP
#define ORACLE_SHIM_BOOST_THREAD_HPP
#include <thread>
#include <mutex>
#include <cstdlib>
namespace boost {
class thread {
    std::thread _t;
  public:
    thread() { }
    template<class F> explicit thread(F f) : _t(f) { }
    ~thread() { if(_t.joinable()) _t.detach(); }
    void join() { if(_t.joinable()) _t.join(); }
    bool joinable() const { return _t.joinable(); }
    // ORACLE_THREADS overrides the detected core count (used for the 1-thread baseline run).
    static unsigned hardware_concurrency() {
        const char* e=std::getenv("ORACLE_THREADS");
        if(e && std::atoi(e)>0) return static_cast<unsigned>(std::atoi(e));
        return std::thread::hardware_concurrency(); }
};
class mutex {
    std::mutex _m;
  public:
    void lock() { _m.lock(); }
    void unlock() { _m.unlock(); }
    bool try_lock() { return _m.try_lock(); }
    class scoped_lock {
        mutex& _mx; bool _owns;
      public:
        explicit scoped_lock(mutex& m) : _mx(m), _owns(true) { _mx.lock(); }
        ~scoped_lock() { if(_owns) _mx.unlock(); }
        void lock() { _mx.lock(); _owns=true; }
        void unlock() { _mx.unlock(); _owns=false; }
    };
};
}
#endif
