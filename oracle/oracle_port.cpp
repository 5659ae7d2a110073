/*
 * oracle/oracle_port.cpp — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * The self-contained CPU oracle: restated runtime (rt_port.h) + restated glue (driver.inc).
 * Parity is PINNED: tests/test_oracle_golden.py checks it against the reference's own
 * known-answer vectors (JoinHashTableTest.cpp, GroupByHashTest.cpp, PartitionedGroupByTest.cpp,
 * test_pyhdk_api.py) and, where /root/reference is present, against oracle/_ref built from
 * the reference's sources.
 */
#include "rt_port.h"
#define ORACLE_KIND "port"
#define ORACLE_KIND_FN oracle_kind
#include "driver.inc"
