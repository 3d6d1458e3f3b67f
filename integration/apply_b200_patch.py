#!/usr/bin/env python3
"""Applies the SRSRAN_B200 hooks of INTEGRATION.md section 2 to the reference's lib/src/phy/phch/sch.c and writes the
result to the path given (a build artefact outside the repository history; nothing of the reference is committed).

    python integration/apply_b200_patch.py /root/reference/lib/src/phy/phch/sch.c /tmp/sch_b200.c
    python integration/apply_b200_patch.py --pdsch /root/reference/lib/src/phy/phch/pdsch.c /tmp/pdsch_b200.c

Three insertions, each keyed on a function signature of the reference:
  decode_tb          -> srsran_b200_decode_tb
  encode_tb_off      -> srsran_b200_encode_tb
  srsran_ulsch_decode-> srsran_b200_ulsch_decode_tb (grants without CQI) / srsran_b200_ulsch_deinterleave
"""
import re
import sys

PROTOS = '''
#ifdef SRSRAN_B200
int srsran_b200_decode_tb(srsran_sch_t* q, srsran_softbuffer_rx_t* softbuffer, srsran_cbsegm_t* cb_segm, uint32_t Qm, uint32_t rv,
                          uint32_t nof_e_bits, int16_t* e_bits, uint8_t* data);
int srsran_b200_takes_tb(srsran_sch_t* q, srsran_cbsegm_t* cb_segm);
int srsran_b200_encode_tb(srsran_sch_t* q, srsran_softbuffer_tx_t* softbuffer, srsran_cbsegm_t* cb_segm, uint32_t Qm, uint32_t rv,
                          uint32_t nof_e_bits, uint8_t* data, uint8_t* e_bits, uint32_t w_offset);
int srsran_b200_ulsch_decode_tb(srsran_sch_t* q, srsran_softbuffer_rx_t* softbuffer, srsran_cbsegm_t* cb_segm, uint32_t Qm, uint32_t rv,
                                int16_t* q_bits, uint32_t H_prime_total, uint32_t N_pusch_symbs, srsran_uci_bit_t* ri_bits,
                                uint32_t nof_ri_bits, uint32_t e_offset, uint32_t nof_e_bits, int16_t* g_bits, uint32_t nof_g_out, uint8_t* data);
int srsran_b200_ulsch_deinterleave(int16_t* q_bits, uint32_t Qm, uint32_t H_prime_total, uint32_t N_pusch_symbs, int16_t* g_bits,
                                   srsran_uci_bit_t* ri_bits, uint32_t nof_ri_bits);
#endif
'''


def insert_after_open_brace(src, signature_regex, text):
    m = re.search(signature_regex, src)
    if not m:
        raise SystemExit("anchor not found: " + signature_regex)
    brace = src.index("{", m.end())
    return src[:brace + 1] + "\n" + text + src[brace + 1:]


PDSCH_PROTOS = '''
#ifdef SRSRAN_B200
uint32_t srsran_b200_pdsch_c_init(uint16_t rnti, int codeword_idx, uint32_t nslot, uint32_t cell_id);
int      srsran_b200_dlsch_takes_symbols(srsran_sch_t* q, srsran_pdsch_cfg_t* cfg, int tb_idx);
int      srsran_b200_dlsch_decode2_symbols(srsran_sch_t* q, srsran_pdsch_cfg_t* cfg, cf_t* symbols, uint32_t c_init, uint8_t* data, int tb_idx,
                                           uint32_t nof_layers);
#endif
'''


def patch_pdsch(src):
    """srsran_pdsch_codeword_decode (pdsch.c:662-760): demodulate + descramble + srsran_dlsch_decode2 become ONE call with the
    equalised symbols when nothing else needs the LLRs on the host (no EVM measurement, no CSI correction, 16-bit mode)"""
    last_inc = [m for m in re.finditer(r'^#include .*$', src, re.M)][-1]
    src = src[:last_inc.end()] + "\n" + PDSCH_PROTOS + src[last_inc.end():]
    m = re.search(r'static int srsran_pdsch_codeword_decode\(', src)
    if not m:
        raise SystemExit("anchor not found: srsran_pdsch_codeword_decode")
    a = src.index("    /* demodulate symbols", m.end())
    b = src.index("    ret = srsran_dlsch_decode2(dl_sch, cfg, q->e[codeword_idx], data[tb_idx].payload, tb_idx, nof_layers);", a)
    b = src.index("\n", b) + 1
    head = '''#ifdef SRSRAN_B200
    if (!(cfg->meas_evm_en && q->evm_buffer[codeword_idx]) && !cfg->csi_enable && srsran_b200_dlsch_takes_symbols(dl_sch, cfg, tb_idx)) {
      data[tb_idx].evm = NAN;
      ret = srsran_b200_dlsch_decode2_symbols(dl_sch, cfg, q->d[codeword_idx],
                                              srsran_b200_pdsch_c_init(cfg->rnti, codeword_idx, 2 * (sf->tti % SRSRAN_NOF_SF_X_FRAME), q->cell.id),
                                              data[tb_idx].payload, tb_idx, nof_layers);
    } else {
#endif
'''
    tail = "#ifdef SRSRAN_B200\n    }\n#endif\n"
    return src[:a] + head + src[a:b] + tail + src[b:]


def main():
    if len(sys.argv) > 3 and sys.argv[1] == "--pdsch":
        open(sys.argv[3], "w").write(patch_pdsch(open(sys.argv[2]).read()))
        return
    src = open(sys.argv[1]).read()
    # prototypes after the last #include
    last_inc = [m for m in re.finditer(r'^#include .*$', src, re.M)][-1]
    src = src[:last_inc.end()] + "\n" + PROTOS + src[last_inc.end():]
    src = insert_after_open_brace(src, r'static int decode_tb\(srsran_sch_t\*\s+q,[^)]*\)',
                                  "#ifdef SRSRAN_B200\n  if (q != NULL && srsran_b200_takes_tb(q, cb_segm)) { /* int16 LLRs, and 8-bit LLRs of sizes with an 8-bit decoder */\n    return srsran_b200_decode_tb(q, softbuffer, cb_segm, Qm, rv, nof_e_bits, e_bits, data);\n  }\n#endif\n")
    src = insert_after_open_brace(src, r'static int encode_tb_off\(srsran_sch_t\*\s+q,[^)]*\)',
                                  "#ifdef SRSRAN_B200\n  return srsran_b200_encode_tb(q, softbuffer, cb_segm, Qm, rv, nof_e_bits, data, e_bits, w_offset);\n#endif\n")
    # srsran_ulsch_decode: right after Q_prime_ri is known
    m = re.search(r'int srsran_ulsch_decode\(', src)
    if not m:
        raise SystemExit("anchor not found: srsran_ulsch_decode")
    a = src.index("uint32_t Q_prime_ri = (uint32_t)ret;", m.end())
    a = src.index("\n", a) + 1
    hook = '''#ifdef SRSRAN_B200
  if (!cfg->uci_cfg.cqi.data_enable && !q->llr_is_8bit) { /* no CQI in this grant: de-interleave + decode in one device submission */
    if (cb_segm.tbs == 0) {
      return ret; /* what the reference returns here: the value left by uci_decode_ri_ack */
    }
    return srsran_b200_ulsch_decode_tb(q, cfg->softbuffers.rx, &cb_segm, Qm, cfg->grant.tb.rv, q_bits, nb_q / Qm, cfg->grant.nof_symb,
                                       q->ack_ri_bits, Q_prime_ri * Qm, 0, (nb_q / Qm - Q_prime_ri) * Qm, NULL, 0, data);
  }
#endif
'''
    src = src[:a] + hook + src[a:]
    open(sys.argv[2], "w").write(src)


if __name__ == "__main__":
    main()
