#include "errors.h"
#include "drvae_b200.h"
namespace drvae {
char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}
}  // namespace drvae
extern "C" const char* drvae_last_error(void) { return drvae::error_buffer(); }
