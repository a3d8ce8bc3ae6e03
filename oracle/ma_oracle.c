#include "ma_oracle.h"
int ma_oracle_dummy;
