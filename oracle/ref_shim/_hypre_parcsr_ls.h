#include "hypre_stub.h"
