#ifndef ORACLE_STUB_PB_H_
#define ORACLE_STUB_PB_H_
#endif
