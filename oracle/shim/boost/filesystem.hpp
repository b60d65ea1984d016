// oracle/shim/boost/filesystem.hpp -- TEST INFRASTRUCTURE: maps the one boost call the
// reference makes (filesystem::exists, external/fworkdir.hpp:20) onto std::filesystem.
#pragma once
#include <filesystem>
namespace boost { namespace filesystem { using std::filesystem::exists; using std::filesystem::path; } }
