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

// loop batches queries from InCh (≈1 M reads or 256 MB), calls the device once per batch and emits
// QueryResults; Query/Seq objects go back to poolQuery/poolSeq exactly as U:337-341 does.
func (e *gpuEngine) loop() {
	defer e.wg.Done()
	const maxQ, maxB = 1 << 20, 256 << 20
	batch := make([]*Query, 0, maxQ)
	var seq []byte        // Go memory is fine: the library does not retain it after the call returns
	off := make([]C.uint64_t, 1, maxQ*2+1)
	flush := func() {
		if len(batch) == 0 {
			return
		}
		var p C.kmcpg_search_params
		C.kmcpg_default_params(&p)
		p.min_query_len, p.min_matched = C.int32_t(e.Options.MinQLen), C.int32_t(e.Options.MinMatched)
		p.dedup_threshold, p.min_query_cov = C.int32_t(e.Options.DeduplicateThreshold), C.double(e.Options.MinQueryCov)
		if batch[0].Seq2 != nil {
			p.paired = 1
		}
		var hits C.kmcpg_hits
		rc := C.kmcpg_search_batch(e.ctx, &p, (*C.uint8_t)(unsafe.Pointer(&seq[0])), &off[0], C.uint32_t(len(off)-1), &hits)
		if rc != 0 {
			checkError(fmt.Errorf("kmcp-gpu: %s", C.GoString(C.kmcpg_last_error(e.ctx)))) // log + os.Exit(-1), util-cli.go:35-40
		}
		nk := unsafe.Slice((*int32)(unsafe.Pointer(hits.n_kmers)), len(batch))
		ql := unsafe.Slice((*int32)(unsafe.Pointer(hits.query_len)), len(batch))
		hs := unsafe.Slice((*C.kmcpg_hit)(unsafe.Pointer(hits.hits)), int(hits.n_hits))
		j := 0
		for q, query := range batch {
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
		batch, seq, off = batch[:0], seq[:0], off[:1]
	}
	for query := range e.InCh {
		batch = append(batch, query)
		seq = append(seq, query.Seq.Seq...)
		off = append(off, C.uint64_t(len(seq)))
		if query.Seq2 != nil {
			seq = append(seq, query.Seq2.Seq...)
			off = append(off, C.uint64_t(len(seq)))
		}
		if len(batch) == maxQ || len(seq) >= maxB {
			flush()
		}
	}
	flush()
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
