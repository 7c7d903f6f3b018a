/* Exceptions.hpp — error types of the SVGDCpp API (reference Exceptions.hpp:16-50): same class names
 * and message prefixes, so user code that catches them keeps working on the accelerated path. */
#ifndef SVGDCPP_B200_EXCEPTIONS_HPP
#define SVGDCPP_B200_EXCEPTIONS_HPP

#include <exception>
#include <stdexcept>
#include <string>

#define SVGDCPP_LOG_PREFIX std::string("SVGDCpp: ")

class DimensionMismatchException : public std::exception {
public:
    explicit DimensionMismatchException(const std::string &what_arg) : text_(SVGDCPP_LOG_PREFIX + "[Dimension Error] " + what_arg) {}
    const char *what() const noexcept override { return text_.c_str(); }

private:
    std::string text_;
};

class UnsetException : public std::exception {
public:
    explicit UnsetException(const std::string &what_arg) : text_(SVGDCPP_LOG_PREFIX + "[Unset Error] " + what_arg) {}
    const char *what() const noexcept override { return text_.c_str(); }

private:
    std::string text_;
};

namespace svgdcpp_b200 {
/* Maps a C-ABI status (include/svgd_b200.h) to the reference's exception convention. */
inline void ThrowOnError(int status, const char *message)
{
    const std::string msg = message ? message : "";
    switch (status) {
    case 0: return;
    case -2: throw DimensionMismatchException(msg);
    case -3: throw UnsetException(msg);
    case -1: throw std::invalid_argument(SVGDCPP_LOG_PREFIX + msg);
    default: throw std::runtime_error(SVGDCPP_LOG_PREFIX + "[Runtime Error] " + msg);
    }
}
} // namespace svgdcpp_b200
#endif
