#include "../../cvstub.hpp"
