// placeholders until each app lands (gl_app_create reports "not available")
#include "app_base.h"
namespace gl {
gl_app* make_cdlp() { return nullptr; }
gl_app* make_lcc() { return nullptr; }
}  // namespace gl
