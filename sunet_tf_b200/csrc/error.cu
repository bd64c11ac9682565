#include "error.h"
namespace sunet {
char* last_error_buf() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace sunet
