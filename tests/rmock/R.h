/* TEST INFRASTRUCTURE: see Rinternals.h in this directory. */
#ifndef RMOCK_R_H
#define RMOCK_R_H
#include <stddef.h>
#endif
