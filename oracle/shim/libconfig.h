/* Stand-in for <libconfig.h> when building the reference from /root/reference
 * (test infrastructure only; see oracle/build_ref.sh, which puts
 * miluphcuda_b200/csrc on the include path). Forwards to the repo's own
 * libconfig-format reader. */
#include "libconfig_lite.h"
