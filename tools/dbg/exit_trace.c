// developer aid: LD_PRELOAD shim that prints a backtrace when the process calls exit() / _exit()
#define _GNU_SOURCE
#include <execinfo.h>
#include <unistd.h>
#include <stdlib.h>
#include <stdio.h>
#include <sys/syscall.h>
static void bt(const char* what, int c) {
  void* b[64];
  int n = backtrace(b, 64);
  fprintf(stderr, "[exit_trace] %s(%d)\n", what, c);
  backtrace_symbols_fd(b, n, 2);
}
void exit(int c) { bt("exit", c); syscall(SYS_exit_group, c); __builtin_unreachable(); }
void _exit(int c) { bt("_exit", c); syscall(SYS_exit_group, c); __builtin_unreachable(); }
