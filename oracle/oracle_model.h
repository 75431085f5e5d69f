// ORACLE (test infrastructure) -- model function table used by the CPU restatement.
// One OracleModel = the generated closures of one workload: the counterparts of the reference's
// Dynamics / Objective / Constraint structs (reference src/dynamics.jl:1-11, src/objectives.jl:1-10,
// src/constraints.jl:1-12).  All matrices are dense, column-major, every entry written.
#pragma once
#include "detmath.h"

typedef struct OracleModel {
  const char* name;
  int nx, nu, nc, nxn, np;
  void (*dyn)(const double* x, const double* u, const double* p, double* f);
  void (*cost)(const double* x, const double* u, const double* p, double* l);
  void (*costN)(const double* x, const double* p, double* l);
  void (*con)(const double* x, const double* u, const double* p, double* c);
  void (*derivs)(const double* x, const double* u, const double* v, const double* p, double* fx, double* fu,
                 double* lx, double* lu, double* lxx, double* luu, double* lux, double* cx, double* cu,
                 double* vcxx, double* vcux, double* vcuu);
  void (*vf)(const double* x, const double* u, const double* v, const double* p, double* vfxx, double* vfux,
             double* vfuu);
  void (*derivsN)(const double* x, const double* p, double* lx, double* lxx);
  int nxt;   // state size costN / derivsN are evaluated on (= nx except for the last stage type of a chain)
} OracleModel;
