// stand-in for <boost/thread.hpp>: the compiled reference files only use boost::mutex with scoped_lock (single-threaded tests)
#ifndef UVIP_BOOST_THREAD_STANDIN
#define UVIP_BOOST_THREAD_STANDIN
namespace boost {
class mutex {
public:
    struct scoped_lock { explicit scoped_lock(mutex&) {} };
};
template <class M> struct unique_lock { explicit unique_lock(M&) {} };
}
#endif
