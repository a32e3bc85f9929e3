/* cdyn stand-in: util.c includes this header but calls nothing from it. */
#ifndef CDYNSHIM_SIMULATE_H
#define CDYNSHIM_SIMULATE_H
#endif
