/* oracle/shim: see RcppArmadillo.h (test infrastructure only) */
#include "RcppArmadillo.h"
