#include "../../include/nfk.h"
extern "C" int nfk_version(void) { return 1; }
