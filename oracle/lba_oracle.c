/* placeholder: filled in with the local-BA restatement */
#include "oracle.h"
