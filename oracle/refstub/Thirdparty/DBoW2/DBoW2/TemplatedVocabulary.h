#include "FORB.h"
