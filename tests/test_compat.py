"""The reference-facing C++ boundary (include/compat/*.h): the reference's own main() (src/BreakID.cc up to the
closing brace of main, compiled at build time by oracle/Makefile `compat_main` against include/compat/BreakID.h and linked
with libbreakid_compat.so) must write the same call files as the reference binary."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
COMPAT_MAIN = os.path.join(ROOT, "oracle", "_ref", "BreakID_compat_main")
COMPAT_LIB = os.path.join(ROOT, "breakid_b200", "host", "libbreakid_compat.so")
STAGE_FUNCTIONS = ["get_mean_insert_size", "scan_discordant_pairs", "add_enspan_point_id", "remove_isolated_pairs", "find_cluster_pairs_enspan_ahc",
                   "find_cluster_pairs_enspan_fast", "findClusterBreakPointInfoSaTag", "write_enspan_out", "write_enspan_params", "annotate_cluster_for_sa_tag",
                   "determine_fusion_type_from_drp", "build_pair_array", "add_cluster_id_for_enspan_vec", "init_cluster", "print_root_nodes",
                   "combine_genome_chr_pos", "get_right_neighbor_sequence_nib", "get_left_neighbor_sequence_nib", "chromID2ChrName",
                   "find_longest_repeat_substring", "split_string"]


def test_compat_headers_compile_on_their_own(tmp_path):
    """every header of include/compat is self-contained (no samtools / htslib on the include path)"""
    inc = os.path.join(ROOT, "include", "compat")
    for h in sorted(os.listdir(inc)):
        src = tmp_path / ("use_" + h + ".cc")
        src.write_text('#include "%s"\nint main() { return 0; }\n' % h)
        subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I" + inc, str(src)])


def test_compat_library_defines_the_reference_stage_functions():
    if not os.path.exists(COMPAT_LIB):
        pytest.skip("libbreakid_compat.so not built (run __graft_entry__.build())")
    out = subprocess.run(["nm", "-DC", "--defined-only", COMPAT_LIB], capture_output=True, text=True, check=True).stdout
    defined = {line.split(" T ", 1)[1].split("(")[0].replace("[abi:cxx11]", "") for line in out.splitlines() if " T " in line}
    missing = [f for f in STAGE_FUNCTIONS if f not in defined]
    assert not missing, missing


def test_neighbour_sequences_and_genome_coordinates(tmp_path):
    """host-only helpers of util_bam.h against their definition (src/util_bam.cc:57-142), through a small C++ program"""
    if not os.path.exists(COMPAT_LIB):
        pytest.skip("libbreakid_compat.so not built")
    import numpy as np
    from breakid_b200 import synth
    n = 1000
    packed = synth.random_nib_bytes(n, 77).numpy()
    nibdir = tmp_path / "nib"
    nibdir.mkdir()
    with open(nibdir / "hg19_chrT.nib", "wb") as f:
        f.write(np.array([0x6be93d3a, n], dtype="<u4").tobytes())
        f.write(packed.tobytes())
    code = "TCAGN"       # nib nibble codes 0..4 (high nibble first); the soft-mask bit is ignored
    seq = "".join(code[min(((packed[i // 2] >> (4 if i % 2 == 0 else 0)) & 7), 4)] for i in range(n))
    src = tmp_path / "t.cc"
    src.write_text('''#include "util_bam.h"
#include <cstdio>
int main(int, char **argv) {
  printf("%s\\n%s\\n%s\\n", get_right_neighbor_sequence_nib("chrT", 100, 21, argv[1]).c_str(), get_left_neighbor_sequence_nib("chrT", 100, 20, argv[1]).c_str(),
         get_sequence_nib("chrT", 5, 14, argv[1]).c_str());
  uint32_t lens[3] = {4000000000u, 500000000u, 7u}; char *names[3] = {0, 0, 0};
  bam_header_t h{3, names, lens};
  printf("%u %u\\n%s %s %s [%s]\\n", combine_genome_chr_pos(&h, 1, 10), combine_genome_chr_pos(&h, 2, 3), chromID2ChrName(0).c_str(), chromID2ChrName(22).c_str(),
         chromID2ChrName(23).c_str(), chromID2ChrName(24).c_str());
  return 0; }
''')
    exe = tmp_path / "t"
    host = os.path.dirname(COMPAT_LIB)
    csrc = os.path.join(ROOT, "breakid_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include", "compat"), str(src), "-o", str(exe), "-L" + host, "-lbreakid_compat",
                           "-Wl,-rpath," + host, "-Wl,-rpath," + csrc, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe), str(nibdir)], capture_output=True, text=True, check=True).stdout.split("\n")
    assert out[0] == seq[100:121]                      # right neighbour of 1-based 100: 0-based [100, 121)
    assert out[1] == seq[79:99]                        # left neighbour: 0-based [79, 99)
    assert out[2] == seq[4:14]
    assert out[3] == "%d %d" % ((4000000000 + 10) % 2**32, (4000000000 + 500000000 + 3) % 2**32)
    assert out[4] == "chr1 chrX chrY []"


def test_string_helpers_of_util_bed(tmp_path):
    """find_longest_repeat_substring against the reference's own function (oracle/_ref/libbreakid_ref.so) and the oracle's
    restatement on random base strings; split_string against its definition"""
    if not os.path.exists(COMPAT_LIB):
        pytest.skip("libbreakid_compat.so not built")
    import numpy as np
    import oracle_py as O
    rng = np.random.RandomState(3)
    cases = ["A", "AC", "AAAAAAAAAAAAC", "ACGT", "TTTTTTTTTTTGGGGGGGGGGGGG"] + ["".join(rng.choice(list("ACGTN"), p=[.4, .2, .2, .15, .05], size=int(rng.randint(1, 60)))) for _ in range(200)]
    src = tmp_path / "t.cc"
    src.write_text('''#include "util_bed.h"
#include <cstdio>
#include <iostream>
int main() {
  std::string s;
  while (std::getline(std::cin, s)) printf("%d\\n", find_longest_repeat_substring(s));
  for (auto &p : split_string("a,,b,c,", ",")) printf("[%s]", p.c_str());
  for (auto &p : split_string("x--y", "--")) printf("[%s]", p.c_str());
  for (auto &p : split_string("whole", "")) printf("[%s]", p.c_str());
  printf("\\n");
  return 0; }
''')
    exe = tmp_path / "t"
    host = os.path.dirname(COMPAT_LIB)
    csrc = os.path.join(ROOT, "breakid_b200", "csrc")
    subprocess.check_call(["g++", "-std=c++17", "-I" + os.path.join(ROOT, "include", "compat"), str(src), "-o", str(exe), "-L" + host, "-lbreakid_compat",
                           "-Wl,-rpath," + host, "-Wl,-rpath," + csrc, "-Wl,-rpath,/usr/local/cuda/lib64"])
    out = subprocess.run([str(exe)], input="\n".join(cases) + "\n", capture_output=True, text=True, check=True).stdout.split("\n")
    got = [int(x) for x in out[:len(cases)]]
    assert got == [O.olib().orc_longest_repeat(c.encode()) for c in cases]
    if O.have_ref():
        assert got == [O.rlib().ref_longest_repeat(c.encode()) for c in cases]
    assert out[len(cases)] == "[a][b][c][x][y][whole]"


@pytest.mark.gpu
@pytest.mark.parametrize("flags", [["-all"], ["-all", "-fast"], ["-q", "30"]])
def test_reference_main_built_against_compat_headers_writes_reference_call_files(tmp_path, flags):
    import oracle_py as O
    from breakid_b200 import bamio, synth
    if not O.have_ref() or not os.path.exists(COMPAT_MAIN):
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    cfg = synth.SynthConfig(chrom_lens=[300000, 200000, 150000], n_tra=3, n_inv=2, n_dup=2, n_del=2, seed=31, sv_jitter=1)
    d = synth.generate(cfg)
    paths = bamio.write_dataset(str(tmp_path), d, genes_per_mb=25.0)
    O.ref_index(paths["bam"])
    O.ref_install_refgene(paths["refgene"])
    ref = subprocess.run([O.REF_BIN, "-i", paths["bam"], "-o", str(tmp_path / "ref"), "-n", paths["nib"]] + flags, capture_output=True, text=True, timeout=600)
    assert ref.returncode == 0, ref.stderr[-1000:]
    env = dict(os.environ, BREAKID_REFGENE=paths["refgene"])
    got = subprocess.run([COMPAT_MAIN, "-i", paths["bam"], "-o", str(tmp_path / "gpu"), "-n", paths["nib"]] + flags, capture_output=True, text=True, timeout=600, env=env)
    assert got.returncode == 0, (got.stdout[-1000:], got.stderr[-1000:])
    suffixes = ["_fusion.txt"] + (["_fusion_all.txt"] if "-all" in flags else [])
    for suffix in suffixes:
        a = open(str(tmp_path / "ref") + suffix).read()
        assert a == open(str(tmp_path / "gpu") + suffix).read(), suffix
        if suffix == "_fusion_all.txt":
            assert len(a.splitlines()) >= 5
    pa = open(str(tmp_path / "ref") + "_params.txt").read().replace(str(tmp_path / "ref"), "X")
    assert pa == open(str(tmp_path / "gpu") + "_params.txt").read().replace(str(tmp_path / "gpu"), "X")
    # the stage messages both mains print from the stage results agree too: insert statistics, distance, per-bucket counts
    pick = lambda txt: [l for l in txt.splitlines() if l.startswith(("the insert size mean", "cluster_dist", "discordant pairs found", "the current number of root cluster"))]
    assert pick(ref.stdout) == pick(got.stdout)
