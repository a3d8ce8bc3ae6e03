/* sys.c -- wall/CPU clocks and the "real*cpu/real" stamp used in the [M::fn::stamp] log lines
 * (reference: sys.c:7-46). */
#include <stdio.h>
#include <sys/resource.h>
#include <sys/time.h>
#include "miniasm_b200.h"

static double t_origin;

double sys_cputime(void)
{
	struct rusage u;
	getrusage(RUSAGE_SELF, &u);
	return (u.ru_utime.tv_sec + u.ru_stime.tv_sec) + 1e-6 * (u.ru_utime.tv_usec + u.ru_stime.tv_usec);
}

double sys_realtime(void)
{
	struct timeval tv;
	gettimeofday(&tv, 0);
	return tv.tv_sec + 1e-6 * tv.tv_usec - t_origin;
}

void sys_init(void)
{
#ifdef __linux__
	struct rlimit lim; /* lift the address-space soft limit as far as allowed (sys.c:22-30) */
	if (getrlimit(RLIMIT_AS, &lim) == 0) { lim.rlim_cur = lim.rlim_max; setrlimit(RLIMIT_AS, &lim); }
#endif
	t_origin = 0;
	t_origin = sys_realtime();
}

const char *sys_timestamp(void)
{
	static char stamp[256];
	double real = sys_realtime(), cpu = sys_cputime();
	snprintf(stamp, sizeof(stamp) - 1, "%.3f*%.2f", real, cpu / real);
	return stamp;
}
