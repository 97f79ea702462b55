// pt_error.h — thread-local error message shared by the translation units of libptcore.
#pragma once
#include <string>
namespace ptc {
void set_error(const std::string& msg);
}
