/* oracle stand-in: see h2y_half_stub.h */
#include "h2y_half_stub.h"
