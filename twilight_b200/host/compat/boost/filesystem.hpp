// host/compat: boost::filesystem / boost::system mapped onto <filesystem>. TEST INFRASTRUCTURE ONLY.
#pragma once
#include <filesystem>
#include <system_error>

namespace boost {
namespace filesystem = std::filesystem;
namespace system {
using error_code = std::error_code;
}
} // namespace boost
