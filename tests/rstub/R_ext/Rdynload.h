#include "../R.h"
