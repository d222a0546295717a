// bfvi_internal.h — shared between the translation units of libbfvi_b200.so (not part of the C ABI).
#pragma once
namespace bfvi {
// records the message bfvi_last_error() returns on this host thread and hands `code` back (bfvi_api.cu)
int report_error(int code, const char* fmt, ...);
}  // namespace bfvi
