#include "lmp_stub.h"
