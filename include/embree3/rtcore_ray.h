/* forwarder: the whole API lives in rtcore.h (see there) */
#include "rtcore.h"
