/* Prints the layout of the iterator / result / scorer-args structs.  Built twice:
 *   -DREFERENCE_HEADERS  against /root/reference (-> tests/golden/ii_abi_layout.txt)
 *   default              against include/ii_b200.h (must print the same table). */
#include <stdio.h>
#include <stddef.h>
#ifdef REFERENCE_HEADERS
#include "redisearch.h"
#include "iterators/iterator_api.h"
#include "index_result_rs.h"
typedef QueryIterator QI;
typedef RSIndexResult IR;
typedef RSIndexStats ST;
#define TAG_METRIC RSResultData_Metric
#define TAG_UNION RSResultData_Union
#define TAG_INTERSECTION RSResultData_Intersection
#define TYPE_UNION IteratorType_Union
#define TYPE_INTERSECT IteratorType_Intersect
#define TYPE_EMPTY IteratorType_Empty
#define TYPE_METRIC_BY_ID IteratorType_MetricSortedById
#else
#include HDR
typedef II_QueryIterator QI;
typedef II_IndexResult IR;
typedef II_IndexStats ST;
#define TAG_METRIC II_ResultData_Metric
#define TAG_UNION II_ResultData_Union
#define TAG_INTERSECTION II_ResultData_Intersection
#define TYPE_UNION II_IteratorType_Union
#define TYPE_INTERSECT II_IteratorType_Intersect
#define TYPE_EMPTY II_IteratorType_Empty
#define TYPE_METRIC_BY_ID II_IteratorType_MetricSortedById
#endif
#define S(n, t) printf("sizeof(" n ")=%zu\n", sizeof(t))
#define O(n, t, f) printf("offsetof(" n "," #f ")=%zu\n", offsetof(t, f))
#define E(n, e) printf(n "=%d\n", (int)(e))
int main(void) {
  S("QueryIterator", QI); O("QueryIterator", QI, type); O("QueryIterator", QI, atEOF); O("QueryIterator", QI, lastDocId);
  O("QueryIterator", QI, current); O("QueryIterator", QI, NumEstimated); O("QueryIterator", QI, Read);
  O("QueryIterator", QI, SkipTo); O("QueryIterator", QI, Revalidate); O("QueryIterator", QI, Free);
  O("QueryIterator", QI, Rewind); O("QueryIterator", QI, ProfileChildren); O("QueryIterator", QI, PrintProfile);
  S("RSIndexResult", IR); O("RSIndexResult", IR, docId); O("RSIndexResult", IR, dmd); O("RSIndexResult", IR, fieldMask);
  O("RSIndexResult", IR, freq); O("RSIndexResult", IR, data); O("RSIndexResult", IR, metrics); O("RSIndexResult", IR, weight);
  O("RSIndexResult", IR, hasFieldExpiration);
  { IR r; printf("offsetof(RSIndexResult,data.metric)=%zu\n", (size_t)((char *)&r.data.metric - (char *)&r)); }
  S("RSIndexStats", ST); O("RSIndexStats", ST, numDocs); O("RSIndexStats", ST, numTerms); O("RSIndexStats", ST, avgDocLen);
  E("ITERATOR_OK", ITERATOR_OK); E("ITERATOR_NOTFOUND", ITERATOR_NOTFOUND); E("ITERATOR_EOF", ITERATOR_EOF);
  E("ITERATOR_TIMEOUT", ITERATOR_TIMEOUT); E("VALIDATE_OK", VALIDATE_OK); E("VALIDATE_MOVED", VALIDATE_MOVED);
  E("VALIDATE_ABORTED", VALIDATE_ABORTED); E("VALIDATE_TIMEOUT", VALIDATE_TIMEOUT);
  E("TAG_METRIC", TAG_METRIC); E("TAG_UNION", TAG_UNION); E("TAG_INTERSECTION", TAG_INTERSECTION);
  E("TYPE_UNION", TYPE_UNION); E("TYPE_INTERSECT", TYPE_INTERSECT); E("TYPE_EMPTY", TYPE_EMPTY);
  E("TYPE_METRIC_BY_ID", TYPE_METRIC_BY_ID);
  return 0;
}
