// CPU emulation of the device-side record parser (matchtigs_b200/csrc/parse.cu): the kernel section of that file is
// compiled as host code behind a few shims and every "thread" runs sequentially; the results are compared with a
// straightforward line-by-line reader on random FASTA / bcalm2 texts (CRLF, multi-line records, blank lines, malformed
// links, wrong ids, non-ACGT characters).  TEST HARNESS ONLY -- built and run by tests/test_parse_emulation.py, which
// extracts parse_kernels.inc from parse.cu; nothing here is linked into the product.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <random>
#include <array>
using u8=uint8_t; using u32=uint32_t; using u64=uint64_t;
#define __global__
#define __device__
#define __forceinline__ inline
#define __launch_bounds__(x)
#define __restrict__
struct uint4 { u32 x,y,z,w; };
struct Idx { unsigned x; } blockIdx, threadIdx;
static inline u32 __vcmpeq4(u32 a,u32 b){ u32 r=0; for(int i=0;i<4;i++) if(((a>>(8*i))&0xFF)==((b>>(8*i))&0xFF)) r|=0xFFu<<(8*i); return r; }
static inline int __ffs(u32 x){ return __builtin_ffs((int)x); }
static inline int __popc(u32 x){ return __builtin_popcount(x); }
static inline int __clz(u32 x){ return x?__builtin_clz(x):32; }
static inline int atomicExch(int* p,int v){ int o=*p; *p=v; return o; }
static inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v){ auto o=*p; *p|=v; return o; }
using std::min;
namespace k {
#include "parse_kernels.inc"
}
using namespace k;
template<class F> void launch(u64 n, F f){ for(u64 i=0;i<n;i++){ blockIdx.x=(unsigned)(i/TB); threadIdx.x=(unsigned)(i%TB); f(); } }

struct Ref { std::vector<std::string> seqs; std::vector<std::array<u64,4>> links; int err=0; };
// straightforward reference reader (same rules as csrc/reader.cpp / the oracle)
Ref ref_parse(const std::string& t, bool bcalm){
    Ref r; size_t i=0, L=t.size(); bool any=false;
    std::vector<std::string> lines; 
    size_t pos=0;
    while(pos<L){ size_t e=t.find('\n',pos); if(e==std::string::npos) e=L; lines.push_back(t.substr(pos,e-pos)); pos=e+1; }
    for(auto& ln: lines){
        if(!ln.empty() && ln[0]=='>'){
            any=true; r.seqs.emplace_back();
            if(bcalm){
                // id
                size_t q=1; u64 id=0; bool d=false; while(q<ln.size()&&isdigit((unsigned char)ln[q])){id=id*10+(ln[q]-'0');q++;d=true;}
                if(!d||id!=r.seqs.size()-1) r.err=2;
                for(size_t p=1;p<ln.size();p++){
                    if(ln[p]=='L' && (ln[p-1]==' '||ln[p-1]=='\t') && p+1<ln.size() && ln[p+1]==':'){
                        // token length >= 7 within the line (and within the text: pos+6 < L)
                        bool ok = p+6 < ln.size()+ (size_t)0 ; // token chars p..p+6 must exist and be non-blank, non-eol
                        if(ok) for(int qn=2;qn<7;qn++){ char c=ln[p+qn]; if(c==' '||c=='\t'||c=='\r') ok=false; }
                        if(!ok) continue;
                        char s=ln[p+2]; size_t c=p+4; u64 n=0; bool dg=false; while(c<ln.size()&&isdigit((unsigned char)ln[c])){n=n*10+(ln[c]-'0');c++;dg=true;}
                        bool shape = ln[p+3]==':' && dg && c+1<ln.size() && ln[c]==':' && ln[c+1]!=' '&&ln[c+1]!='\t'&&ln[c+1]!='\r';
                        if(!shape){ r.err=3; r.links.push_back(std::array<u64,4>{(u64)r.seqs.size()-1,0,0,0}); continue; }
                        char tt=ln[c+1];
                        if((s!='+'&&s!='-')||(tt!='+'&&tt!='-')) r.err=4;
                        r.links.push_back(std::array<u64,4>{(u64)r.seqs.size()-1,(u64)(s==0x2b),n,(u64)(tt==0x2b)});
                    }
                }
            }
        } else {
            for(char c: ln){ if(c=='\r') continue; if(!any){ r.err=1; } else r.seqs.back().push_back(c); }
        }
    }
    return r;
}
int run_case(const std::string& text, bool bcalm, bool verbose){
    u64 L=text.size(); const char* d=text.data();
    u64 nch=(L+CHUNK-1)/CHUNK; bool aligned = ((uintptr_t)d&15)==0;
    std::vector<u32> key(nch+1), packed(nch+1), nrec(nch+1), nseq(nch+1), nlink(nch+1), rb(nch+1), sb(nch+1), lb(nch+1);
    launch((nch+TB-1)/TB*TB, [&]{ chunk_scan_text(d,L,nch,aligned,bcalm,key.data(),packed.data()); });
    for(u64 j=1;j<nch;j++) key[j]=std::max(key[j],key[j-1]);
    launch((nch+TB-1)/TB*TB, [&]{ resolve_counts(key.data(),packed.data(),nch,nrec.data(),nseq.data(),nlink.data()); });
    u64 U=0,B=0,NL=0; for(u64 j=0;j<nch;j++){ rb[j]=U; U+=nrec[j]; sb[j]=B; B+=nseq[j]; lb[j]=NL; NL+=nlink[j]; }
    if(!bcalm) NL=0;
    std::vector<unsigned long long> words((B+31)/32+2,0); std::vector<u64> off(U+1), la(NL+1), lbb(NL+1); std::vector<u8> sa(NL+1), sbb(NL+1); int err=0;
    launch((nch+TB-1)/TB*TB, [&]{ chunk_scatter(d,L,nch,aligned,key.data(),bcalm,rb.data(),sb.data(),lb.data(),words.data(),off.data(),la.data(),sa.data(),lbb.data(),sbb.data(),&err); });
    off[U]=B;
    Ref r=ref_parse(text,bcalm);
    auto fail=[&](const char* m){ if(verbose) printf("FAIL %s (bcalm=%d, L=%zu) err=%d ref.err=%d\n",m,(int)bcalm,(size_t)L,err,r.err); return 1; };
    if(r.err==1){ return err==1?0:fail("expected err1"); }
    if(err==1) return fail("unexpected err1");
    if(U!=r.seqs.size()) return fail("record count");
    // sequences (only compare when all ACGT)
    bool acgt=true; for(auto&s:r.seqs) for(char c:s) if(c!='A'&&c!='C'&&c!='G'&&c!='T') acgt=false;
    if(!acgt){ if(err!=6 && !(r.err && err==r.err)) return fail("expected err6"); return 0; }
    if(err==6) return fail("unexpected err6");
    u64 p=0;
    for(u64 u=0;u<U;u++){
        if(off[u]!=p) return fail("offset");
        for(char c: r.seqs[u]){ u32 code=(words[p>>5]>>(2*(p&31)))&3; if("ACTG"[code]!=c) return fail("base"); p++; }
    }
    if(p!=B) return fail("total bases");
    if(bcalm){
        if(r.err==2){ return err==2?0: (err==3||err==4)?0:fail("expected err2"); }
        if(NL!=r.links.size()) return fail("link count");
        if(r.err){ return err?0:fail("expected link err"); }
        if(err) return fail("unexpected err");
        for(u64 i=0;i<NL;i++) if(la[i]!=r.links[i][0]||sa[i]!=r.links[i][1]||lbb[i]!=r.links[i][2]||sbb[i]!=r.links[i][3]) return fail("link");
    } else if(err) return fail("unexpected err (fasta)");
    return 0;
}
int main(int argc, char** argv){
    const int cases = argc > 1 ? atoi(argv[1]) : 20000;
    std::mt19937_64 rng(7); int bad=0, n=0;
    for(int it=0; it<cases && bad<5; it++){
        bool bcalm = rng()&1;
        int nrec = rng()%6; std::string t;
        if(rng()%8==0) t += std::string(rng()%3,'\n');
        bool crlf = rng()%5==0;
        std::string nl = crlf? "\r\n":"\n";
        for(int r=0;r<nrec;r++){
            t += ">";
            if(bcalm){ t += std::to_string((rng()%20==0)? r+1 : r); } else { t += "rec"+std::to_string(rng()%1000); }
            int nf = rng()%4;
            for(int f=0;f<nf;f++){
                t += (rng()&1)?" ":"\t";
                int kind=rng()%6;
                if(kind<3){ t += "L:"; t += "+-x"[rng()%20==0?2:rng()%2]; t += ":"; t += std::to_string(rng()%30000); t += ":"; t += "+-"[rng()%2]; }
                else if(kind==3) t += "LN:i:"+std::to_string(rng()%100);
                else if(kind==4) t += "L:+:"+std::string(rng()%2,'7');  // short / malformed
                else t += "km:f:1.0";
            }
            t += nl;
            int nlines = 1 + rng()%3;
            for(int l=0;l<nlines;l++){ int len=rng()%90; for(int i=0;i<len;i++) t += (rng()%3000==0)?'N':"ACGT"[rng()%4]; if(l+1<nlines || r+1<nrec || rng()%6) t += nl; }
        }
        n++;
        if(run_case(t,bcalm,true)){ bad++; printf("---\n%s\n---\n", t.c_str()); }
    }
    printf("%d cases, %d failures\n", n, bad);
    return bad!=0;
}
