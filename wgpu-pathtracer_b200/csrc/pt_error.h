// pt_error.h — thread-local error message shared by the translation units of libptcore, and the guard that keeps C++
// exceptions (std::bad_alloc from a host vector, std::system_error from a thread that cannot start) inside the library:
// the C ABI reports a negative status and a message, it never throws or aborts across the boundary.
#pragma once
#include <exception>
#include <new>
#include <string>
namespace ptc {
void set_error(const std::string& msg);

// runs `body` (a callable returning a ptc_status int); an escaping exception becomes -4 (PTC_ERR_OOM) or -1 (PTC_ERR_INVALID)
template <class F>
int guarded(const char* where, F&& body) noexcept {
    try { return body(); }
    catch (const std::bad_alloc&) { try { set_error(std::string(where) + ": out of host memory"); } catch (...) {} return -4; }
    catch (const std::exception& e) { try { set_error(std::string(where) + ": " + e.what()); } catch (...) {} return -1; }
    catch (...) { try { set_error(std::string(where) + ": unknown internal error"); } catch (...) {} return -1; }
}
}
