// TEST INFRASTRUCTURE: boost::lexical_cast<T>(std::string) for the ParameterReader (include/parameter_reader.h:52-61).
#ifndef SSM_REFSTUB_BOOST_LEXICAL_CAST
#define SSM_REFSTUB_BOOST_LEXICAL_CAST
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
template <typename T> inline T lexical_cast(const std::string& s)
{
    std::istringstream in(s);
    T v;
    if (!(in >> v)) throw std::runtime_error("bad lexical cast: " + s);
    return v;
}
template <> inline std::string lexical_cast<std::string>(const std::string& s) { return s; }
}  // namespace boost
#endif
