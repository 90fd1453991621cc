//! Drop into the reference's `basic/tests/` (see rust/parity/README.md of valida_b200).  Proves the Fibonacci program with the
//! reference's own prover under a DETERMINISTIC Poseidon instance and compares the CBOR bytes with the digest valida_b200 records
//! for the same program and constants.  Run with RAYON_NUM_THREADS=1 (smallest proof-of-work witness).
//!
//! NOT COMPILED in valida_b200's development container (no Rust toolchain there).

extern crate alloc;

use std::fs;
use std::path::PathBuf;

use p3_baby_bear::BabyBear;
use p3_challenger::DuplexChallenger;
use p3_dft::Radix2Bowers;
use p3_field::extension::BinomialExtensionField;
use p3_field::{AbstractField, Field};
use p3_fri::{FriConfig, TwoAdicFriPcs, TwoAdicFriPcsConfig};
use p3_keccak::Keccak256Hash;
use p3_mds::coset_mds::CosetMds;
use p3_merkle_tree::FieldMerkleTreeMmcs;
use p3_poseidon::Poseidon;
use p3_symmetric::{CompressionFunctionFromHasher, SerializingHasher32};
use valida_basic::BasicMachine;
use valida_cpu::MachineWithCpuChip;
use valida_machine::{
    FixedAdviceProvider, InstructionWord, Machine, Operands, ProgramROM, StarkConfigImpl,
};
use valida_machine::__internal::p3_commit::ExtensionMmcs;      // as basic/tests/test_prover.rs:33 imports it
use valida_program::MachineWithProgramChip;

type Val = BabyBear;
const P: u64 = 2013265921;

/// The stand-in for the caller's RNG that valida_b200's tests and bench use: SplitMix64 seeded with ASCII "valida",
/// candidates `z >> 33`, kept when below p; 480 canonical words.
fn round_constants() -> Vec<Val> {
    let mut state: u64 = 0x76616c696461;
    let mut out = Vec::with_capacity(480);
    while out.len() < 480 {
        state = state.wrapping_add(0x9E3779B97F4A7C15);
        let mut z = state;
        z = (z ^ (z >> 30)).wrapping_mul(0xBF58476D1CE4E5B9);
        z = (z ^ (z >> 27)).wrapping_mul(0x94D049BB133111EB);
        z ^= z >> 31;
        let cand = z >> 33;
        if cand < P {
            out.push(Val::from_canonical_u32(cand as u32));
        }
    }
    out
}

fn load_program(dir: &PathBuf) -> Vec<InstructionWord<i32>> {
    fs::read_to_string(dir.join("fib25.words"))
        .expect("fib25.words (python rust/parity/make_inputs.py <dir>)")
        .lines()
        .filter(|l| !l.trim().is_empty())
        .map(|l| {
            let w: Vec<i32> = l.split_whitespace().map(|x| x.parse().unwrap()).collect();
            assert_eq!(w.len(), 6);
            InstructionWord { opcode: w[0] as u32, operands: Operands([w[1], w[2], w[3], w[4], w[5]]) }
        })
        .collect()
}

#[test]
fn pin_parity() {
    let dir = PathBuf::from(std::env::var("PARITY_DIR").expect("PARITY_DIR"));
    let program = load_program(&dir);

    // the machine exactly as prove_program sets it up (basic/tests/test_prover.rs:403-411)
    let mut machine = BasicMachine::<Val>::default();
    let rom = ProgramROM::new(program);
    machine.program_mut().set_program_rom(&rom);
    machine.cpu_mut().fp = 0x1000;
    machine.cpu_mut().save_register_state();
    machine.run(&rom, &mut FixedAdviceProvider::empty());
    assert_eq!(machine.cpu().clock, 192);

    // the configuration of the reference's own tests (test_prover.rs:412-455), Poseidon constants fixed
    type Challenge = BinomialExtensionField<Val, 5>;
    type PackedChallenge = BinomialExtensionField<<Val as Field>::Packing, 5>;
    type Mds16 = CosetMds<Val, 16>;
    type Perm16 = Poseidon<Val, Mds16, 16, 5>;
    type MyHash = SerializingHasher32<Keccak256Hash>;
    type MyCompress = CompressionFunctionFromHasher<Val, MyHash, 2, 8>;
    type ValMmcs = FieldMerkleTreeMmcs<Val, MyHash, MyCompress, 8>;
    type ChallengeMmcs = ExtensionMmcs<Val, Challenge, ValMmcs>;
    type Dft = Radix2Bowers;
    type Challenger = DuplexChallenger<Val, Perm16, 16>;
    type MyFriConfig = TwoAdicFriPcsConfig<Val, Challenge, Challenger, Dft, ValMmcs, ChallengeMmcs>;
    type Pcs = TwoAdicFriPcs<MyFriConfig>;
    type MyConfig = StarkConfigImpl<Val, Challenge, PackedChallenge, Pcs, Challenger>;

    let perm16 = Perm16::new(4, 22, round_constants(), Mds16::default());
    let hash = MyHash::new(Keccak256Hash {});
    let compress = MyCompress::new(hash);
    let val_mmcs = ValMmcs::new(hash, compress);
    let challenge_mmcs = ChallengeMmcs::new(val_mmcs.clone());
    let fri_config = FriConfig { log_blowup: 1, num_queries: 40, proof_of_work_bits: 8, mmcs: challenge_mmcs };
    let pcs = Pcs::new(fri_config, Dft::default(), val_mmcs);
    let config = MyConfig::new(pcs, Challenger::new(perm16));

    let proof = machine.prove(&config);
    machine.verify(&config, &proof).expect("the reference's own verifier");
    let mut bytes = vec![];
    ciborium::into_writer(&proof, &mut bytes).expect("serialization");
    fs::write(dir.join("reference_proof.cbor"), &bytes).unwrap();

    let expected = fs::read_to_string(dir.join("expected.json")).expect("expected.json");
    let want_len: usize = field(&expected, "proof_bytes").parse().unwrap();
    println!("reference proof: {} bytes (valida_b200: {})", bytes.len(), want_len);
    println!("compare: sha256sum {}/reference_proof.cbor   against   {}", dir.display(), field(&expected, "sha256"));
    assert_eq!(bytes.len(), want_len, "proof sizes differ: a shape convention (CBOR, query count, tree arity) is read differently");
}

/// Minimal field extraction from the small JSON file (no serde_json in the reference's dev-dependencies).
fn field(json: &str, key: &str) -> String {
    let at = json.find(&format!("\"{}\"", key)).expect(key);
    let rest = &json[at + key.len() + 2..];
    let rest = &rest[rest.find(':').unwrap() + 1..];
    rest.trim_start().trim_start_matches('"').split(|c| c == '"' || c == ',' || c == '\n' || c == '}').next().unwrap().trim().to_string()
}
