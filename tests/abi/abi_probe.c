#include <stdio.h>
#include <stddef.h>
#include HDR
#define S(t) printf("sizeof(" #t ")=%zu\n", sizeof(t))
#define O(t,f) printf("offsetof(" #t "," #f ")=%zu\n", offsetof(t,f))
#define E(e) printf(#e "=%d\n", (int)(e))
int main(void){
 S(VecSimParams); S(AlgoParams); S(BFParams); S(HNSWParams); S(SVSParams); S(TieredIndexParams);
 S(VecSimQueryParams); S(VecSimIndexBasicInfo); S(VecSimIndexStatsInfo); S(VecSimRawParam); S(VecSim_InfoField); S(VecSimMemoryFunctions);
 S(VecSimIndexDebugInfo); S(CommonInfo); S(tieredInfoStruct); S(svsInfoStruct); S(hnswInfoStruct);
 O(VecSimIndexDebugInfo,commonInfo); O(VecSimIndexDebugInfo,bfInfo); O(VecSimIndexDebugInfo,tieredInfo);
 O(CommonInfo,basicInfo); O(CommonInfo,indexSize); O(CommonInfo,indexLabelCount); O(CommonInfo,memory); O(CommonInfo,lastMode);
 O(tieredInfoStruct,backendCommonInfo); O(tieredInfoStruct,frontendCommonInfo); O(tieredInfoStruct,bfInfo); O(tieredInfoStruct,management_layer_memory); O(tieredInfoStruct,backgroundIndexing); O(tieredInfoStruct,bufferLimit);
 O(VecSimParams,algo); O(VecSimParams,algoParams); O(VecSimParams,logCtx);
 O(BFParams,type); O(BFParams,dim); O(BFParams,metric); O(BFParams,multi); O(BFParams,initialCapacity); O(BFParams,blockSize);
 O(VecSimQueryParams,batchSize); O(VecSimQueryParams,searchMode); O(VecSimQueryParams,timeoutCtx);
 O(VecSimIndexBasicInfo,algo); O(VecSimIndexBasicInfo,metric); O(VecSimIndexBasicInfo,type); O(VecSimIndexBasicInfo,isMulti); O(VecSimIndexBasicInfo,isTiered); O(VecSimIndexBasicInfo,isDisk); O(VecSimIndexBasicInfo,blockSize); O(VecSimIndexBasicInfo,dim);
 O(VecSimIndexStatsInfo,memory); O(VecSimIndexStatsInfo,numberOfMarkedDeleted); O(VecSimIndexStatsInfo,directHNSWInsertions); O(VecSimIndexStatsInfo,flatBufferSize);
 O(VecSimRawParam,name); O(VecSimRawParam,nameLen); O(VecSimRawParam,value); O(VecSimRawParam,valLen);
 O(VecSim_InfoField,fieldName); O(VecSim_InfoField,fieldType); O(VecSim_InfoField,fieldValue);
 E(VecSimType_FLOAT32); E(VecSimType_FLOAT64); E(VecSimType_BFLOAT16); E(VecSimType_FLOAT16); E(VecSimType_INT8); E(VecSimType_UINT8); E(VecSimType_INT32); E(VecSimType_INT64);
 E(VecSimAlgo_BF); E(VecSimAlgo_HNSWLIB); E(VecSimAlgo_TIERED); E(VecSimAlgo_SVS);
 E(VecSimMetric_L2); E(VecSimMetric_IP); E(VecSimMetric_Cosine);
 E(BY_SCORE); E(BY_ID); E(BY_SCORE_THEN_ID); E(VecSim_QueryReply_OK); E(VecSim_QueryReply_TimedOut);
 E(VecSimParamResolver_OK); E(VecSimParamResolverErr_NullParam); E(VecSimParamResolverErr_AlreadySet); E(VecSimParamResolverErr_UnknownParam); E(VecSimParamResolverErr_BadValue);
 E(VecSimParamResolverErr_InvalidPolicy_NExits); E(VecSimParamResolverErr_InvalidPolicy_NHybrid); E(VecSimParamResolverErr_InvalidPolicy_NRange); E(VecSimParamResolverErr_InvalidPolicy_AdHoc_With_BatchSize); E(VecSimParamResolverErr_InvalidPolicy_AdHoc_With_EfRuntime);
 E(EMPTY_MODE); E(STANDARD_KNN); E(HYBRID_ADHOC_BF); E(HYBRID_BATCHES); E(HYBRID_BATCHES_TO_ADHOC_BF); E(RANGE_QUERY);
 E(QUERY_TYPE_NONE); E(QUERY_TYPE_KNN); E(QUERY_TYPE_HYBRID); E(QUERY_TYPE_RANGE);
 E(INFOFIELD_STRING); E(INFOFIELD_INT64); E(INFOFIELD_UINT64); E(INFOFIELD_FLOAT64); E(INFOFIELD_ITERATOR);
 return 0; }
