// TEST INFRASTRUCTURE: boost::timer (src/mapper.cpp:111,162) and the boost smart-pointer names PCL's typedefs use.
#ifndef SSM_REFSTUB_BOOST_TIMER
#define SSM_REFSTUB_BOOST_TIMER
#include <chrono>
#include <memory>
namespace boost {
class timer {
public:
    timer() : t0_(std::chrono::steady_clock::now()) {}
    double elapsed() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
private:
    std::chrono::steady_clock::time_point t0_;
};
template <typename T> using shared_ptr = std::shared_ptr<T>;
template <typename T, typename... A> inline std::shared_ptr<T> make_shared(A&&... a) { return std::make_shared<T>(static_cast<A&&>(a)...); }
}  // namespace boost
#endif
