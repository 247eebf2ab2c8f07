// engine_gpu.go — the cgo binding of libkmcp_gpu.so for the reference (drop into kmcp/cmd/, build with `-tags gpu`, CGO_ENABLED=1).
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no Go toolchain.  This file is the stub of INTEGRATION.md as a source file,
// kept in step with include/kmcp_gpu.h by hand; the same C symbols are exercised by the C++ CLI, the ctypes binding and the C99
// examples, which are what the tests run.

//go:build gpu

package cmd

/*
#cgo CFLAGS: -I${SRCDIR}/../../third_party/kmcp_b200/include
#cgo LDFLAGS: -L${SRCDIR}/../../third_party/kmcp_b200 -lkmcp_gpu
#include <stdlib.h>
#include "kmcp_gpu.h"
*/
import "C"

import (
	"errors"
	"fmt"
	"sync"
	"unsafe"
)

// gpuEngine offers the same five touch points as UnikIndexDBSearchEngine (U:192-203).
type gpuEngine struct {
	Options SearchOptions
	DBs     []*UnikIndexDB // Info only (search.go:405-409, 790 read DBs[0].Info)
	InCh    chan *Query
	OutCh   chan *QueryResult
	ctx     *C.kmcpg_ctx
	info    C.kmcpg_db_info_t
	queryFPR func(n, k int) float64
	wg      sync.WaitGroup
}

func newGPUEngine(opt SearchOptions, device int, dbPath string) (*gpuEngine, error) {
	e := &gpuEngine{Options: opt}
	if rc := C.kmcpg_create(C.int(device), &e.ctx); rc != 0 {
		return nil, fmt.Errorf("kmcp-gpu: %s", C.GoString(C.kmcpg_last_error(nil)))
	}
	p := C.CString(dbPath)
	defer C.free(unsafe.Pointer(p))
	if rc := C.kmcpg_open_db(e.ctx, p, nil); rc != 0 {
		return nil, fmt.Errorf("open kmcp db: %s: %s", dbPath, C.GoString(C.kmcpg_last_error(e.ctx)))
	}
	C.kmcpg_db_info(e.ctx, &e.info)
	info, err := UnikIndexDBInfoFromFile(filepath.Join(dbPath, dbInfoFile)) // unchanged Go code
	if err != nil {
		return nil, err
	}
	e.DBs = []*UnikIndexDB{{Info: info}}
	e.queryFPR = QueryFPRWithCacheWithConstantFPR(opt.FPRBufSize, float64(e.info.fpr)) // util-fpr.go:140
	e.InCh = make(chan *Query, 1<<16)
	e.OutCh = make(chan *QueryResult, 1<<16)
	e.wg.Add(1)
	go e.loop()
	return e, nil
}

// One batch under construction or in flight.  seq and off live in PINNED C memory (kmcpg_host_alloc): kmcpg_search_submit keeps
// reading them until kmcpg_search_wait returns, and cgo forbids C code to hold on to Go memory after a call has returned.
type gpuBatch struct {
	queries []*Query
	seq     unsafe.Pointer // maxB bytes
	off     unsafe.Pointer // (2*maxQ+1) uint64
	nSeq    int
	nBytes  int
	job     *C.kmcpg_job
}

const maxQ, maxB = 1 << 20, 256 << 20

func newGpuBatch() *gpuBatch {
	b := &gpuBatch{queries: make([]*Query, 0, maxQ)}
	C.kmcpg_host_alloc(&b.seq, C.size_t(maxB+1<<16))
	C.kmcpg_host_alloc(&b.off, C.size_t(8*(2*maxQ+1)))
	return b
}

func (b *gpuBatch) add(s []byte) {
	copy(unsafe.Slice((*byte)(unsafe.Add(b.seq, b.nBytes)), len(s)), s)
	b.nBytes += len(s)
	b.nSeq++
	unsafe.Slice((*C.uint64_t)(b.off), 2*maxQ+1)[b.nSeq] = C.uint64_t(b.nBytes)
}

// loop batches queries from InCh (≈1 M reads or 256 MB) and keeps TWO batches submitted (kmcpg_search_submit): while the device
// works on batch i, batch i+1 is already queued behind it — the executor starts its first part the moment the last part of batch i
// leaves the compute stream — and this goroutine turns the hits of batch i-1 into QueryResults.  Query/Seq objects go back to
// poolQuery/poolSeq exactly as U:337-341 does.
func (e *gpuEngine) loop() {
	defer e.wg.Done()
	free := []*gpuBatch{newGpuBatch(), newGpuBatch(), newGpuBatch()}
	var inflight []*gpuBatch
	cur := free[0]
	free = free[1:]
	submit := func(b *gpuBatch) {
		var p C.kmcpg_search_params
		C.kmcpg_default_params(&p)
		p.min_query_len, p.min_matched = C.int32_t(e.Options.MinQLen), C.int32_t(e.Options.MinMatched)
		p.dedup_threshold, p.min_query_cov = C.int32_t(e.Options.DeduplicateThreshold), C.double(e.Options.MinQueryCov)
		if b.queries[0].Seq2 != nil {
			p.paired = 1
		}
		var bt C.kmcpg_batch
		bt.seq, bt.off, bt.n_seqs = (*C.uint8_t)(b.seq), (*C.uint64_t)(b.off), C.uint32_t(b.nSeq)
		if rc := C.kmcpg_search_submit(e.ctx, &p, &bt, &b.job); rc != 0 {
			checkError(fmt.Errorf("kmcp-gpu: %s", C.GoString(C.kmcpg_last_error(e.ctx)))) // log + os.Exit(-1), util-cli.go:35-40
		}
		inflight = append(inflight, b)
	}
	finish := func() { // the oldest batch in flight
		b := inflight[0]
		inflight = inflight[1:]
		var hits C.kmcpg_hits
		if rc := C.kmcpg_search_wait(b.job, &hits); rc != 0 {
			checkError(fmt.Errorf("kmcp-gpu: %s", C.GoString(C.kmcpg_last_error(e.ctx))))
		}
		nk := unsafe.Slice((*int32)(unsafe.Pointer(hits.n_kmers)), len(b.queries))
		ql := unsafe.Slice((*int32)(unsafe.Pointer(hits.query_len)), len(b.queries))
		hs := unsafe.Slice((*C.kmcpg_hit)(unsafe.Pointer(hits.hits)), int(hits.n_hits))
		j := 0
		for q, query := range b.queries {
			r := poolQueryResult.Get().(*QueryResult)
			r.QueryIdx, r.QueryID, r.QueryLen = query.Idx, query.ID, int(ql[q])
			r.K, r.NumKmers, r.Matches = int(e.info.ks[0]), int(nk[q]), nil
			for ; j < len(hs) && int(hs[j].query) == q; j++ {
				m := e.match(int(nk[q]), hs[j]) // tCov / queryFPR / QCov / Jaccard exactly as U:7466-7491
				if m == nil {
					continue
				}
				if r.Matches == nil {
					r.Matches = poolMatches.Get().(*[]*Match)
				}
				*r.Matches = append(*r.Matches, m)
			}
			sortAndTrim(r, e.Options) // U:273-311, existing code moved into a function
			e.OutCh <- r
			poolSeq.Put(query.Seq)
			if query.Seq2 != nil {
				poolSeq.Put(query.Seq2)
			}
			poolQuery.Put(query)
		}
		C.kmcpg_free_hits(&hits)
		b.queries, b.nSeq, b.nBytes = b.queries[:0], 0, 0
		free = append(free, b)
	}
	flush := func() {
		if len(cur.queries) == 0 {
			return
		}
		submit(cur)
		if len(inflight) == 2 { // two on the device: digest the older one while the newer one runs
			finish()
		}
		cur = free[0]
		free = free[1:]
	}
	for query := range e.InCh {
		cur.queries = append(cur.queries, query)
		cur.add(query.Seq.Seq)
		if query.Seq2 != nil {
			cur.add(query.Seq2.Seq)
		}
		if len(cur.queries) == maxQ || cur.nBytes >= maxB {
			flush()
		}
	}
	flush()
	for len(inflight) > 0 {
		finish()
	}
	close(e.OutCh)
}

func (e *gpuEngine) Wait()        { e.wg.Wait() }                      // U:584-588
func (e *gpuEngine) Close() error { C.kmcpg_close(e.ctx); return nil } // U:591-620

// ---- the reader stage (kmcpg_reader_*) ------------------------------------------------------------------------------------

// in engine_gpu.go: feeds the batches of the input files to the GPU engine instead of sg.InCh <- query
func (e *gpuEngine) searchFiles(read1, read2 string, files []string, wholeFile bool, k int, emit func(*C.kmcpg_read_batch, *C.kmcpg_hits)) error {
	var o C.kmcpg_reader_opts
	C.kmcpg_default_reader_opts(&o)
	o.k = C.int32_t(k)
	if wholeFile { o.whole_file = 1 }
	if read1 != "" {
		o.read1, o.read2 = C.CString(read1), C.CString(read2)          // freed below
		defer C.free(unsafe.Pointer(o.read1)); defer C.free(unsafe.Pointer(o.read2))
	} else {
		arr := C.malloc(C.size_t(len(files)) * C.size_t(unsafe.Sizeof(uintptr(0))))
		defer C.free(arr)
		ptrs := unsafe.Slice((**C.char)(arr), len(files))
		for i, f := range files { ptrs[i] = C.CString(f); defer C.free(unsafe.Pointer(ptrs[i])) }
		o.files, o.n_files = (**C.char)(arr), C.int32_t(len(files))
	}
	var rd *C.kmcpg_reader
	if rc := C.kmcpg_reader_open(&o, &rd); rc != 0 { return fmt.Errorf("kmcpg_reader_open: %d", int(rc)) }
	defer C.kmcpg_reader_close(rd)
	for {
		var b C.kmcpg_read_batch
		switch rc := C.kmcpg_reader_next(rd, &b); {
		case rc == 0:
			return nil
		case rc < 0:
			return errors.New(C.GoString(C.kmcpg_reader_error(rd)))       // search.go: checkError(err)
		}
		var h C.kmcpg_hits
		if rc := C.kmcpg_search_batch(e.ctx, &e.params, b.seq, b.off, b.n_seqs, &h); rc != 0 {
			C.kmcpg_reader_free_batch(&b)
			return errors.New(C.GoString(C.kmcpg_last_error(e.ctx)))
		}
		emit(&b, &h)                                                      // tCov / FPR / sort / TSV as today, Query.Idx = b.first_query + q
		C.kmcpg_free_hits(&h)
		C.kmcpg_reader_free_batch(&b)
	}
}
