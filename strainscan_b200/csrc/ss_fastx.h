// ss_fastx.h -- read-file dialects Jellyfish accepts beyond 4-line FASTQ: FASTA, and FASTQ whose sequence /
// quality are wrapped over several lines (its parser joins the lines: windows spanning a line break count).
// The kernels scan 4-line FASTQ only, so such inputs are rewritten on the host, record by record, into
// 4-line FASTQ with the same sequence characters in the same order (FASTA records get a constant quality
// line).  Semantics follow the reference engine's parser as pinned by the golden vectors jf_wrapped_fastq and
// jf_fasta_reads (tests/golden): file type from the first non-blank byte; FASTA sequence
// = the lines up to the next line opening with '>'; FASTQ sequence = the lines up to a line opening with '+',
// quality = as many following lines as it takes to match the sequence length.  Host only.
#pragma once
#include <stddef.h>
#include <string.h>

#include <vector>

// What to do with a read file, judged by its head (`whole` = the head is the whole file):
//   0 plain 4-line FASTQ (or empty): scan as is;  1 FASTA / wrapped FASTQ / leading blank lines: normalise;
//  -1 not a sequence file
inline int ss_fastx_kind(const char *head, size_t n, bool whole) {
    size_t p = 0;
    while (p < n && (head[p] == '\n' || head[p] == '\r')) p++;
    if (p == n) return 0;
    if (head[p] == '>') return 1;
    if (head[p] != '@') {
        size_t q = p;
        while (q < n && (head[q] == '\n' || head[q] == '\r' || head[q] == ' ' || head[q] == '\t')) q++;
        return (q == n && whole) ? 0 : -1;                      // only whitespace: an empty input
    }
    if (p > 0) return 1;                                        // leading blank lines shift the 4-line framing
    const char *l1 = (const char *)memchr(head, '\n', n);
    if (!l1) return 0;
    const char *l2 = (const char *)memchr(l1 + 1, '\n', (size_t)(head + n - (l1 + 1)));
    if (!l2 || l2 + 1 >= head + n) return 0;                    // a lone record: the framing check of the scan decides
    return l2[1] == '+' ? 0 : 1;
}

// Rewrite FASTA / multi-line FASTQ as 4-line FASTQ.  Returns 0, or -1 for a malformed file (a record that does
// not open with the file's record character, a FASTQ record without '+' line or with a quality of another
// length than its sequence).
inline int ss_fastx_normalize(const char *buf, size_t len, std::vector<char> &out) {
    const char *p = buf, *end = buf + len;
    auto line = [&](const char *&s, size_t &n) {                // the line without its '\n' (a '\r' stays, as std::getline)
        const char *nl = (const char *)memchr(p, '\n', (size_t)(end - p));
        s = p;
        if (nl) { n = (size_t)(nl - p); p = nl + 1; }
        else { n = (size_t)(end - p); p = end; }
    };
    while (p < end && (*p == '\n' || *p == '\r')) p++;
    if (p == end) return 0;
    const char type = *p;
    if (type != '>' && type != '@') return -1;
    out.reserve(out.size() + len + len / 2);
    const char *s;
    size_t n;
    while (p < end) {
        if (*p != type) {                                       // trailing blank lines only
            const char *q = p;
            while (q < end && (*q == '\n' || *q == '\r')) q++;
            return q == end ? 0 : -1;
        }
        line(s, n);
        out.push_back('@');
        out.insert(out.end(), s + 1, s + n);
        out.push_back('\n');
        size_t seq_at = out.size();
        if (type == '>') {
            while (p < end && *p != '>') { line(s, n); out.insert(out.end(), s, s + n); }
            size_t seqlen = out.size() - seq_at;
            out.push_back('\n'); out.push_back('+'); out.push_back('\n');
            out.insert(out.end(), seqlen, 'I');
            out.push_back('\n');
        } else {
            while (p < end && *p != '+') { line(s, n); out.insert(out.end(), s, s + n); }
            if (p == end) return -1;                            // truncated record
            size_t seqlen = out.size() - seq_at;
            line(s, n);                                         // the '+' line
            out.push_back('\n'); out.push_back('+'); out.push_back('\n');
            size_t q = 0;
            while (q < seqlen && p < end) { line(s, n); out.insert(out.end(), s, s + n); q += n; }
            if (q != seqlen) return -1;
            out.push_back('\n');
        }
    }
    return 0;
}
