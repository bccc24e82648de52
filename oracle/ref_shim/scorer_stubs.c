/* TEST INFRASTRUCTURE ONLY.  Symbols src/ext/default.c references from its query EXPANDERS (stemmer,
 * phonetic, synonym, Chinese tokenizer).  The scorer harness never reaches them; they abort if called.
 * Kept in a TU that includes no reference header so the placeholder prototypes cannot clash. */
#include <stdio.h>
#include <stdlib.h>

#define ABORT_STUB(name) \
  void name(void) { fprintf(stderr, "scorer_harness: unexpected call to " #name "\n"); abort(); }
ABORT_STUB(IndexSpec_CheckPhoneticEnabled)
ABORT_STUB(IndexSpec_GetFieldByBit)
ABORT_STUB(NewChineseTokenizer)
ABORT_STUB(NewQueryNode)
ABORT_STUB(PhoneticManager_ExpandPhonetics)
ABORT_STUB(QueryError_SetError)
ABORT_STUB(QueryNode_AddChild)
ABORT_STUB(RSLanguage_ToSnowballStemmer)
ABORT_STUB(SynonymMap_GetIdsBySynonym)
ABORT_STUB(Vector_Free)
ABORT_STUB(__newVectorSize)
ABORT_STUB(__vector_PushPtr)
ABORT_STUB(sb_stemmer_new)
ABORT_STUB(sb_stemmer_delete)
ABORT_STUB(sb_stemmer_stem)
ABORT_STUB(sb_stemmer_length)


/* src/util/mempool/mempool.c reads RSGlobalConfig.noMemPool and logs through RedisModule_Log when REDISEARCH_NO_MEMPOOL is set:
 * a zeroed configuration (pools enabled) and no logger */
char RSGlobalConfig[1 << 16];
void *RedisModule_Log;
