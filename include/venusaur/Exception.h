// Exception.h -- error convention of the drop-in classes; mirrors the reference's Core/Exception.h:90-154
// (class Exception : std::runtime_error thrown by the *_CHECK macros).  OPTIX_CHECK / GL_CHECK have no equivalent
// here: there is no OptiX and no GL in this path.
#pragma once

#include <sstream>
#include <stdexcept>
#include <string>

#include "../venusaur_b200.h"

namespace venusaur {

class Exception : public std::runtime_error {
public:
    explicit Exception(const char* msg) : std::runtime_error(msg) {}
    explicit Exception(const std::string& msg) : std::runtime_error(msg) {}
};

}  // namespace venusaur

// The C ABI never throws; the shim converts a non-zero status into the reference's exception type, with the call
// text and file:line like CUDA_CHECK (Exception.h:90-102).
#define VN_CHECK(handle, call)                                                                        \
    do {                                                                                              \
        const int vn_status_ = (call);                                                                \
        if (vn_status_ != VN_OK) {                                                                    \
            std::stringstream vn_ss_;                                                                 \
            vn_ss_ << "venusaur_b200 call (" << #call << ") failed with status " << vn_status_ << ": '" \
                   << vn_last_error(handle) << "' (" __FILE__ << ":" << __LINE__ << ")\n";            \
            throw ::venusaur::Exception(vn_ss_.str());                                                \
        }                                                                                             \
    } while (0)
#define VN_MULTI_CHECK(handle, call)                                                                  \
    do {                                                                                              \
        const int vn_status_ = (call);                                                                \
        if (vn_status_ != VN_OK) {                                                                    \
            std::stringstream vn_ss_;                                                                 \
            vn_ss_ << "venusaur_b200 call (" << #call << ") failed with status " << vn_status_ << ": '" \
                   << vn_multi_last_error(handle) << "' (" __FILE__ << ":" << __LINE__ << ")\n";      \
            throw ::venusaur::Exception(vn_ss_.str());                                                \
        }                                                                                             \
    } while (0)
