#include <cuda_runtime.h>
#include <cstdio>
__global__ void k_set(cudaGraphConditionalHandle h, int *cnt, int limit) {
    *cnt += 1;
    cudaGraphSetConditional(h, *cnt < limit);
}
int main() {
    cudaStream_t s; cudaStreamCreate(&s);
    int *cnt; cudaMalloc(&cnt, 4); cudaMemset(cnt, 0, 4);
    cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
    cudaStreamCaptureStatus st; cudaGraph_t g; const cudaGraphNode_t *deps; size_t nd;
    cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &g, &deps, &nd);
    cudaGraphConditionalHandle h;
    cudaGraphConditionalHandleCreate(&h, g, 1, cudaGraphCondAssignDefault);
    cudaGraphNodeParams p = {cudaGraphNodeTypeConditional};
    p.conditional.handle = h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    cudaStreamGetCaptureInfo_v2(s, &st, nullptr, &g, &deps, &nd);
    printf("add %d\n", (int)cudaGraphAddNode(&node, g, deps, nd, &p));
    cudaStreamUpdateCaptureDependencies(s, &node, 1, cudaStreamSetCaptureDependencies);
    cudaGraph_t body = p.conditional.phGraph_out[0];
    cudaStream_t b; cudaStreamCreate(&b);
    cudaStreamBeginCaptureToGraph(b, body, nullptr, nullptr, 0, cudaStreamCaptureModeThreadLocal);
    k_set<<<1, 1, 0, b>>>(h, cnt, 5);
    cudaStreamEndCapture(b, nullptr);
    cudaGraph_t graph; cudaStreamEndCapture(s, &graph);
    cudaGraphExec_t e; printf("inst %d\n", (int)cudaGraphInstantiate(&e, graph, 0));
    cudaGraphLaunch(e, s); cudaStreamSynchronize(s);
    int hc; cudaMemcpy(&hc, cnt, 4, cudaMemcpyDeviceToHost); printf("count %d (expect 5) err %d\n", hc, (int)cudaGetLastError());
}
